"""`model` package shim: the three hot-path modules of the reference's model/ package, backed by awr_b200."""
