/* Hardware probes used while designing the tensor-core kernels (tools/dbg_umma_*.py).  NOT part of the product ABI: they are built into a
 * separate libawr_b200_debug.so (`make -C awr-adaptive-weighting-regression_b200 debug`) that nothing under the package imports. */
#ifndef AWR_B200_DEBUG_H
#define AWR_B200_DEBUG_H
#ifdef __cplusplus
extern "C" {
#endif

/* D[128][64] fp32 = A * Bm^T where A is the row-shifted window {r0 + (m/8)*sbo_rows + m%8} of the TMA-loaded smem tile G[rows][64]
 * (bf16, SWIZZLE_128B); probes UMMA descriptor start-address / SBO / base-offset semantics (tools/dbg_umma_window.py). */
/* MMA issue/execute rate: out_dev[grid] cycles for iters*4 tcgen05.mma (M=128,N,K=16) over nacc accumulators (tools/dbg_umma_rate.py). */
int awr_debug_umma_rate(unsigned long long* out_dev, int N, int nacc, int iters, int a_rows_shift, int grid, void* stream);
/* MMA-issuer loop probe: commit / barrier-wait / TMA-fed ring overheads per group of `per` MMAs (tools/dbg_umma_rate.py). */
int awr_debug_umma_pipe(const void* G, int g_rows, unsigned long long* out_dev, int N, int per, int groups, int mode, int stages,
                        int tma_rows, int grid, void* stream);
int awr_debug_umma_window(const void* G, const void* Bm, float* D, int rows, int r0, int sbo_rows, int base_mode, void* stream);

/* MMA rate with the halo-tile kernel's operand pattern: A window `a_start` bytes into the tile, 8-row groups `sbo` bytes apart, taps
 * `tap_step` bytes apart, two accumulators per weight tile; out_dev[grid] = cycles for iters*8 MMAs (tools/dbg_umma_rate2.py). */
int awr_debug_umma_rate2(unsigned long long* out_dev, int N, int sbo, int a_start, int tap_step, int iters, int grid, void* stream);

/* TMA throughput: every CTA streams `groups` boxes of rows x 128 B through a ring of `stages` buffers; cta_stride_rows = 0: all CTAs read the
 * SAME rows (weight tiles), > 0: private rows; span_rows = how many distinct boxes a CTA cycles through (tools/dbg_tma_rate.py). */
int awr_debug_tma_rate(const void* G, int g_rows, unsigned long long* out_dev, int rows, int stages, int groups, int cta_stride_rows,
                       int span_rows, int mode /* 0 one issuing thread, 1 two threads, 2 two tensor maps, 3 1-D bulk copies */, int grid, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* AWR_B200_DEBUG_H */
