/* cdl_b200.h — C ABI of libcdl_b200.so: the CDLNet K-iteration ISTA forward pass on B200 (sm_100a).
 *
 * The reference (RQLuo/CDLNet-video) has no FFI layer: its hot path is the body of
 *   CDLNet.forward        model/net.py:76-92
 *   CDLNetVideo.forward   model/net.py:192-212
 *   GDLNet.forward        model/net.py:659-675
 * plus pre_process[_3d] / post_process[_3d] (model/utils.py:5-33, 70-98).  This header is the
 * boundary a maintainer binds instead (ctypes stub in INTEGRATION.md): plain pointers and sizes,
 * no torch types.  All data pointers are DEVICE pointers to contiguous fp32 unless a name ends in
 * `_host`.  The library never allocates or frees caller memory and never synchronises the device;
 * work is enqueued on the `stream` argument (a cudaStream_t passed as void*).
 *
 * Return convention: 0 = OK; <0 = argument / shape / unsupported-configuration error (see the
 * enum); >0 = a cudaError_t passed through + CDL_CUDA_ERROR_BASE.  Nothing is printed, nothing
 * throws or aborts.
 */
#ifndef CDL_B200_H
#define CDL_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CDL_ABI_VERSION 1
#define CDL_CUDA_ERROR_BASE 1000
#define CDL_NCCL_ERROR_BASE 2000   /* + ncclResult_t */

enum cdl_status {
  CDL_OK = 0,
  CDL_ERR_NULL = -1,        /* a required pointer is NULL                                   */
  CDL_ERR_SHAPE = -2,       /* non-positive / inconsistent extents                          */
  CDL_ERR_UNSUPPORTED = -3, /* configuration has no kernel (even P, M > 256, ...)           */
  CDL_ERR_ALIGN = -4,       /* pointer not 16-byte aligned                                  */
  CDL_ERR_NO_WEIGHTS = -5,  /* cdl_set_weights has not been called                          */
  CDL_ERR_NO_DEVICE = -6,   /* no CUDA device / not an sm_100 device                        */
  CDL_ERR_RANGE = -7,       /* layer index or slab range out of bounds                      */
  CDL_ERR_WORKSPACE = -8,   /* workspace NULL or too small                                  */
  CDL_ERR_NO_NCCL = -9      /* libnccl.so.2 not loadable (only the multi-GPU entry points need it) */
};

enum cdl_precision {
  CDL_PREC_FP32 = 0, /* CUDA-core fp32 FMA: same arithmetic class as the reference on CPU  */
  CDL_PREC_TF32X3 = 2, /* like TF32 with the ANALYSIS convolution as a 3-term split, r_hi W_hi + r_lo W_hi + r_hi W_lo (fp32-class
                      * accuracy of the step that carries the error at large K, ~1.6x the analysis time).  2-D stride-1
                      * geometries only; elsewhere the FP32 kernels run.                                              */
  CDL_PREC_TF32 = 1  /* tcgen05 kind::tf32, operands rounded RNE, fp32 accumulate in TMEM; the final D z as a 3-term (hi/lo) split.
                      * Covered geometries: the video network (3-D, 7x7x7, s = 2, C = 1, M <= 176, model/net.py:123-143)
                      * and the 2-D stride-1 networks (7x7, s = 1, C <= 3, M <= 64, padded width % 4 == 0:
                      * model/net.py:20-36 with args.json / JDD args, GDLNet :572-600); any other geometry gets the
                      * FP32 kernels - see cdl_plan_precision().                                                      */
};

/* Geometry of one plan.  Mirrors the constructor kwargs of the reference modules
 * (model/net.py:20-28, 123-133, 572-582) plus the input extents.
 *
 * Temporal slabs (multi-GPU): a rank owning coarse frames [q0,q1) of a longer clip
 * creates its plan with dims[0] = the number of fine frames it keeps resident (its own s*(q1-q0)
 * frames plus `halo_front` frames before and `halo_back` after), and sets halo_front/halo_back
 * (Pd/2 and Pd/2-s+1 at a seam, 0 at a true clip end).  With both 0 the plan is the ordinary
 * unsharded operator with zero padding Pd/2 at both ends.                                        */
typedef struct cdl_desc {
  int32_t ndim;        /* 2 (images N,C,H,W) or 3 (clips N,C,D,H,W)                          */
  int32_t N, C, M, K;  /* batch, image channels, subbands, unrolled iterations              */
  int32_t dims[3];     /* UNPADDED input extents (D,H,W); D ignored (1) when ndim == 2       */
  int32_t P[3];        /* filter extents (Pd,Ph,Pw), odd; Pd ignored (1) when ndim == 2      */
  int32_t s;           /* stride                                                             */
  int32_t has_mask;    /* 1: a mask tensor shaped like y multiplies B z (JDD, net.py:87)     */
  int32_t precision;   /* enum cdl_precision                                                 */
  int32_t halo_front;  /* temporal slab: extra fine frames held before the owned range       */
  int32_t halo_back;   /* temporal slab: extra fine frames held after the owned range        */
  int32_t device;      /* CUDA device ordinal                                                */
} cdl_desc_t;

typedef struct cdl_plan cdl_plan_t;

/* Derived index layout, bit-exact with calc_pad_2D / calc_pad_3D (model/utils.py:35-51,100-108). */
typedef struct cdl_layout {
  int32_t pad[6];      /* (left,right,top,bottom,front,back) — the reference's tuple order   */
  int32_t fine[3];     /* padded (D,H,W) of yp / residual / xphat                            */
  int32_t coarse[3];   /* (D,H,W) of the sparse code z                                       */
} cdl_layout_t;

int         cdl_abi_version(void);
const char* cdl_status_string(int status);

int  cdl_plan_create(cdl_plan_t** out, const cdl_desc_t* desc);
void cdl_plan_destroy(cdl_plan_t* plan);
int  cdl_plan_layout(const cdl_plan_t* plan, cdl_layout_t* out);
int  cdl_plan_workspace_bytes(const cdl_plan_t* plan, size_t* out);
/* The workspace regions are ordered so that smaller uses are prefixes of the full workspace: a driver that only calls
 * cdl_reduce_sums / cdl_mean_from_sums allocates cdl_plan_reduce_workspace_bytes, one that only calls the stepwise entry
 * points (cdl_analysis_step, cdl_synthesis_step, cdl_preprocess) cdl_plan_step_workspace_bytes.                      */
int  cdl_plan_reduce_workspace_bytes(const cdl_plan_t* plan, size_t* out);
int  cdl_plan_step_workspace_bytes(const cdl_plan_t* plan, size_t* out);
/* Effective precision after plan creation (a TF32 request falls back to FP32 kernels for
 * geometries the tensor-core path does not cover; never to the CPU).                          */
int  cdl_plan_precision(const cdl_plan_t* plan);

/* Filters and thresholds.  A[k], B[k]: K device pointers (HOST array of pointers) to the
 * (M,C,Pd,Ph,Pw) weights of nn.Conv / nn.ConvTranspose (model/net.py:32-33,137-142) or to the
 * synthesised Gabor filters (model/gabor.py:46-51).  t: (K,2,M) thresholds (model/net.py:35,144).
 * Repacks into the kernels' layouts; call again after any weight change.                        */
int cdl_set_weights(cdl_plan_t* plan, const float* const* A, const float* const* B, const float* t, void* stream);

/* pre_process / pre_process_3d (model/utils.py:5-22, 70-87).
 * sums: optional (2N) doubles out: per-sample sum(y), sum(mask) (or element count).             */
int cdl_reduce_sums(cdl_plan_t* plan, const float* y, const float* mask, double* sums, void* workspace, void* stream);
int cdl_mean_from_sums(cdl_plan_t* plan, const double* sums, float* mean, void* stream);
int cdl_center_pad(cdl_plan_t* plan, const float* y, const float* mask, const float* mean, float* yp, float* mask_p, void* stream);
int cdl_preprocess(cdl_plan_t* plan, const float* y, const float* mask, float* yp, float* mask_p, float* mean, void* workspace, void* stream);

/* The stepwise entry points keep the sparse code in the plan's INTERNAL layout (`code`, cdl_plan_code_bytes bytes):
 * (N,M,coarse) for the fp32 kernels and the 2-D tensor-core kernels, an opaque blocked channels-last layout (176 padded
 * subbands, rows padded to 8 sites) for the video tensor-core kernels - treat it as a byte buffer.  Convert with cdl_code_export / cdl_code_import; cdl_forward / cdl_denoise always return z as (N,M,coarse).                      */
int cdl_plan_code_bytes(const cdl_plan_t* plan, size_t* out);
int cdl_code_export(cdl_plan_t* plan, const float* code, float* z, void* stream);
int cdl_code_import(cdl_plan_t* plan, const float* z, float* code, void* stream);

/* One analysis step  code <- ST(code -/+ A_k r, t[k,0] + c*t[k,1])   (model/net.py:85,87;200,205).
 * first != 0: code <- ST(A_k r, .) with r = yp (iteration 0, no input code).  In place.  c: N floats or NULL. */
int cdl_analysis_step(cdl_plan_t* plan, int k, int first, const float* r, const float* c, float* code, void* workspace, void* stream);
/* One synthesis step  out <- mask_p * B_k z - yp  (residual != 0) or out <- B_k z (residual == 0). */
int cdl_synthesis_step(cdl_plan_t* plan, int k, int residual, const float* code, const float* yp, const float* mask_p, float* out, void* workspace, void* stream);

/* Opt-in for stepwise drivers that alternate cdl_synthesis_step(residual, out = R) and cdl_analysis_step(r = R) on the
 * SAME buffer R with the same yp (the loop of model/net.py:204-205 unrolled by the caller, e.g. the temporal-slab
 * driver): the analysis step may then overwrite R - its input, dead once read - with -yp, and the next residual
 * synthesis into R skips its own initialisation pass (one launch and one image pass less per iteration).  Only the
 * tensor-core path uses it; cdl_forward always does this on its own buffer.  enable = 0 restores const semantics.    */
int cdl_plan_set_rearm(cdl_plan_t* plan, int enable);

/* All K iterations + D z:  z (N,M,coarse) and xphat (N,C,fine) out.  z may be NULL: the code then stays in the
 * workspace and the pass that converts it to (N,M,coarse) is skipped.                             */
int cdl_forward(cdl_plan_t* plan, const float* yp, const float* mask_p, const float* c, float* z, float* xphat, void* workspace, void* stream);
/* post_process / post_process_3d (model/utils.py:24-33, 89-98): crop the stride padding, add the mean. */
int cdl_postprocess(cdl_plan_t* plan, const float* xphat, const float* mean, float* xhat, void* stream);

/* pre + forward + post on device buffers: y, mask (N,C,dims) -> xhat (N,C,dims), z (N,M,coarse).   */
int cdl_denoise(cdl_plan_t* plan, const float* y, const float* mask, const float* c, float* xhat, float* z, void* workspace, void* stream);
/* Same with HOST buffers (pinned for true asynchrony): copies y/mask/c in, xhat (and z if z_host
 * is non-NULL) out, on `stream`.  The device staging area lives in `workspace` (see
 * cdl_plan_host_workspace_bytes).  This is the end-to-end entry bench.py's `e2e` times.            */
int cdl_plan_host_workspace_bytes(const cdl_plan_t* plan, size_t* out);
int cdl_plan_host_workspace_bytes_noz(const cdl_plan_t* plan, size_t* out);   /* enough when z_host == NULL */
int cdl_denoise_host(cdl_plan_t* plan, const float* y_host, const float* mask_host, const float* c_host,
                     float* xhat_host, float* z_host, void* workspace, void* stream);

/* ---- temporal slabs across GPUs (the reference has no counterpart: analyze3d.py:62,105-106 chops a
 * clip into independent 16-frame windows instead).  One process per GPU; rank r's plan is a slab plan (halo_front /
 * halo_back above).  Per iteration the ranks exchange the Pd - s seam frames of their partial B z with their ring
 * neighbours (NCCL P2P send/recv, grouped) and add them; the sum is fused into the analysis step's rounding pass.
 * NCCL is dlopen'ed (libnccl.so.2) on first use.                                                                     */
typedef struct cdl_comm cdl_comm_t;
int  cdl_comm_unique_id(void* id128);                         /* rank 0: 128-byte ncclUniqueId to hand to every rank */
int  cdl_comm_create(cdl_comm_t** out, const void* id128, int rank, int nranks, int device);
void cdl_comm_destroy(cdl_comm_t* comm);
int  cdl_comm_allreduce_f64(cdl_comm_t* comm, double* buf, size_t n, void* stream);   /* the global mean's sums */
int  cdl_halo_bytes(const cdl_plan_t* plan, size_t* out);     /* bytes of ONE receive buffer: (N, Pd - s, Fh, Fw) fp32 */
int  cdl_halo_exchange(cdl_plan_t* plan, cdl_comm_t* comm, const float* r, float* recv_prev, float* recv_next, void* stream);
/* analysis step on r + received seam partials (+ yp back on the seams when yp != NULL: residual mode, both partials
 * carry -yp); recv_* as written by cdl_halo_exchange, ignored on a side without a halo                                */
int  cdl_analysis_step_halo(cdl_plan_t* plan, int k, const float* r, const float* c, float* code, const float* recv_prev,
                            const float* recv_next, const float* yp, void* workspace, void* stream);
int  cdl_halo_add(cdl_plan_t* plan, float* r, const float* recv_prev, const float* recv_next, const float* yp, void* stream);
/* All K iterations + D z of one rank's slab; r (N,C,Fd,Fh,Fw) returns xphat on the resident frames; halo_ws holds two
 * receive buffers (2 x cdl_halo_bytes); workspace >= cdl_plan_step_workspace_bytes.  comm may be NULL without halos.  */
int  cdl_forward_sharded(cdl_plan_t* plan, cdl_comm_t* comm, const float* yp, const float* c, float* code, float* r,
                         void* halo_ws, void* workspace, void* stream);

/* ---- frame-recurrent CSR variants.  The analysis step with prox_CSR / prox_CSR_f2 (model/net.py:229-262)
 * in place of ST, as CDLNet_CSR.forward (:426-462) and CDLNet_CSRf2.forward (:525-567) apply it: z_prev / z_after are the
 * neighbouring frames' codes in z's layout (either may be NULL: one neighbour = prox_CSR with that neighbour's gamma, both
 * = prox_CSR_f2, none = the plain step); g1 / g2 = the (K,2,M) gamma parameters paired with z_prev / z_after
 * (gamma = g[k,0] + c*g[k,1]).  Exact fp32 kernels only (plans created with CDL_PREC_FP32): CDL_ERR_UNSUPPORTED otherwise. */
int cdl_analysis_step_csr(cdl_plan_t* plan, int k, int first, const float* r, const float* c, float* code, const float* z_prev,
                          const float* z_after, const float* g1, const float* g2, void* workspace, void* stream);

/* ---- input pipeline fused with pre_process.  Replaces, for a clean clip x on the device,
 *     mask  = utils.gen_bayer_mask(x) (bayer != 0; 2-D, C = 3) | a tensor (mask != NULL) | 1        utils.py:13-27
 *     noisy = mask * (x + noise * (sigma/255))          utils.py:29-55 awgn / awgn3d, analyze3d.py:108-114, train.py:80
 *     yp, mean, mask_p = pre_process[_3d](noisy, s, mask)                                           model/utils.py:5-22,70-87
 * in two passes over x (sums, then centre + pad) instead of ~10.  noise = the caller's randn_like draw (NULL: none),
 * c[n] = sigma_n / 255; y_out (optional) receives the noisy masked clip.  The plan must have been created with
 * has_mask = (bayer || mask).  Roundings follow the reference expression, so yp / mean equal cdl_preprocess on the
 * materialised noisy clip bit for bit.                                                                                */
int cdl_preprocess_noisy(cdl_plan_t* plan, const float* x, const float* noise, const float* c, const float* mask, int bayer,
                         float* y_out, float* yp, float* mask_p, float* mean, void* workspace, void* stream);

/* ---- blind noise level on the device.  Replaces model/nle.py:17-27 `nle_mad` (call sites analyze.py:
 * `255 * model.nle.noise_level(noisy, method=blind)`, analyze3d.py:118-121):
 *     sigma_hat[n] = median(|HH * y[n]|) / 0.6745,   HH = diagonal bior4.4 detail filter (model/wvlt.py:5-42), stride 2,
 * no padding, per channel; lower median over all C x Ho x Wo coefficients (torch.median).  y is (N,C,H,W) fp32 on the
 * device, H, W >= 10; sigma_hat (N) stays on the device (the modules take sigma as a tensor).  Not bound to a plan.    */
int cdl_nle_mad_workspace_bytes(int N, int C, int H, int W, size_t* out);
int cdl_nle_mad(const float* y, int N, int C, int H, int W, float* sigma_hat, void* workspace, void* stream);

/* Number of kernels launched by this plan since creation (bench.py's gpu_launches).               */
int cdl_plan_launch_count(const cdl_plan_t* plan, uint64_t* out);

#ifdef __cplusplus
}
#endif
#endif /* CDL_B200_H */
