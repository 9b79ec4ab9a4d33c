/* flipped-vqa-b200 — test / tuning hooks of libfvqa.so. NOT part of the product boundary (include/fvqa.h): nothing on the
 * training / validation path calls these, so the defaults are never changed there. The knobs are process-wide atomics (a
 * step's backward is launched from autograd's thread, which a thread-local knob would not reach): a test that flips one must
 * not run kernels concurrently from another thread and restores the previous value it gets back. */
#ifndef FVQA_DEBUG_H_
#define FVQA_DEBUG_H_

#ifdef __cplusplus
extern "C" {
#endif

/* 2x2-cluster variant of the GEMM (plain epilogue, M > 128, N % 512 == 0): clusters of TWO CTA pairs own 256 x 512 output blocks
 * and TMA-multicast the A slice the pairs share (-25 % L2 -> SM operand bytes); same results bit for bit. Mode 0 = never,
 * 1 (default) = when the CTA-pair schedule would end in a partial wave (the N = 4096 GEMMs of a 3072-row step), 2 = always when
 * eligible. fvqa_gemm_quad_clusters() = co-resident 4-CTA clusters on this device (33 on a B200), 0 if unavailable. */
int fvqa_gemm_debug_quad(int mode);
int fvqa_gemm_quad_clusters(void);

/* Tuning hook: the skinny (M <= 16) kernel's CTA covers 8 * nt output columns, nt in {1, 2, 4}; 0 = heuristic. */
int fvqa_gemm_debug_skinny_nt(int nt);

/* Test / tuning hook for the GEMM tile choice: bn = multiple of 16 in [64,256] forces that CTA-pair
 * tile width, 0 restores the heuristic, -1 forces the single-CTA kernel. Returns the previous value. */
int fvqa_gemm_debug_force_bn(int bn);

/* Test / tuning hook: epilogue warps per CTA of the CTA-pair kernel: 4 (default) or 8 (two per TMEM lane quadrant). */
int fvqa_gemm_debug_epilogue_warps(int n);

/* Tuning hook: 1 = the CTA-pair kernel's TMA loads carry L2 eviction hints (A evict_last, B evict_first). */
int fvqa_gemm_debug_l2_hints(int on);

/* Probe: 1 = the CTA-pair kernel reads its A operand in the OTHER 16-bit format than B (mixed fp16 x bf16 kind::f16 MMA).
 * Measured on B200: illegal instruction (tools/mixed_umma_probe.py) - which is why the operand format is per build. */
int fvqa_gemm_debug_mixed_a(int on);

/* Tuning hook: 1 (default) = the hot kernels (GEMMs, attention, norms) are launched with programmatic dependent launch, so the next
 * kernel's prologue overlaps the previous kernel's last wave; 0 = ordinary stream-ordered launches. FVQA_PDL=0 in the environment sets
 * 0 at fvqa_init(). Returns the previous setting. */
int fvqa_debug_pdl(int on);

/* Test hook: 0 forces the mma.sync kernels, 1 (default) lets S <= 128, hd = 128 take the tcgen05 path. */
int fvqa_attn_debug_use_tc(int on);

#ifdef __cplusplus
}
#endif
#endif /* FVQA_DEBUG_H_ */
