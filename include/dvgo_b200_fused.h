/*
 * dvgo_b200_fused.h -- C ABI of the FUSED B200 path of libdvgo_b200.so.
 *
 * The op-by-op entry points of dvgo_b200.h reproduce the reference's extension surface one call at
 * a time.  The entry points here implement the same mathematics as the model-level hot path
 *     DirectVoxGO.forward            lib/dvgo.py:450-577   (sample_ray :425-448, mask cascade :469-494,
 *                                    grid_sampler :312-328, compositing :554-576)
 *     loss                           run.py:377-386
 *     backward of all of the above   (autograd in the reference)
 *     TV + MaskedAdam                run.py:389-397, lib/cuda/total_variation_kernel.cu:13-67,
 *                                    lib/cuda/adam_upd_kernel.cu:8-58
 * in a handful of kernels with no host synchronisation and no boolean-mask compaction passes:
 *
 *   ray_setup     per-ray slab test + step count + exclusive scan -> slot offsets (M0 layout)
 *   march_fwd     ONE WARP PER RAY: points, bbox + occupancy cull, density trilinear, alpha, alpha
 *                 threshold, shuffle-scan transmittance with early stop, weight threshold, k0 trilinear
 *                 (channel-last grid: one corner = C contiguous floats), compacted survivor stream
 *   rgb           rgbnet on the survivor stream (tcgen05 kernel, fused_mlp.cu) or sigmoid(k0)
 *   composite     segmented (per-ray) sums of w*rgb and w*step
 *   ray_loss      per-ray loss terms and dL/d(rgb_marched), dL/d(alphainv_last)
 *   sample_grad   per-survivor dL/d(rgb), dL/d(weight)
 *   march_bwd     ONE WARP PER RAY, far to near: alpha2weight + raw2alpha backward, trilinear scatter
 *                 of density and k0 gradients
 *   sweep         TV gradient + (masked) Adam + gradient re-zeroing in one pass over each grid
 *
 * Data layout in HBM
 *   density          [X,Y,Z] fp32          (identical to the reference's [1,1,X,Y,Z])
 *   k0 (channel-last)[X,Y,Z,C] fp32        (the reference keeps [1,C,X,Y,Z]; convert with
 *                                           dvgo_grid_ncdhw_to_cl / dvgo_grid_cl_to_ncdhw at the
 *                                           state_dict boundary)
 *   "slot" arrays    indexed by s = ray_off[r] + step, s < M0 = sum of N_steps: alpha, T, exp_d, code
 *                    code: >= 0 index into the survivor stream; -1 in the transmittance scan but below
 *                    the weight threshold; -2 culled (outside bbox / free space / alpha threshold /
 *                    after the early stop)
 *   survivor stream  compacted, warp-chunk order: feat [M4,C], ray [M4], slot [M4], weight [M4]
 *
 * All pointers are device pointers; conventions as in dvgo_b200.h.
 */
#ifndef DVGO_B200_FUSED_H_
#define DVGO_B200_FUSED_H_

#include <stdint.h>

#include "dvgo_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

/* Scene + render constants shared by the fused kernels (host struct, device pointers inside). */
typedef struct dvgo_scene {
  int X, Y, Z;               /* density / k0 grid size */
  int C;                     /* k0 channels (3, 9 or 12) */
  const float* xyz_min;      /* [3] */
  const float* xyz_max;      /* [3] */
  const uint8_t* mask;       /* occupancy grid [mx,my,mz] bool (lib/dvgo.py:583-613) */
  int mx, my, mz;
  const float* mask_scale;   /* [3] xyz2ijk_scale */
  const float* mask_shift;   /* [3] xyz2ijk_shift */
  float near, far;           /* render_kwargs near / far */
  float stepdist;            /* stepsize * voxel_size (lib/dvgo.py:439) */
  float act_shift;           /* density bias (lib/dvgo.py:61) */
  float interval;            /* stepsize * voxel_size_ratio (lib/dvgo.py:466) */
  float fast_color_thres;    /* alpha / weight threshold (lib/dvgo.py:478,488); 0 disables both masks */
  int ndc;                   /* 0: sample_pts_on_rays sampler; 1: NDC fixed-count sampler (dmpigo) */
  int ndc_samples;           /* N_samples of the NDC sampler (lib/dmpigo.py:188) */
  int exact_transmittance;   /* march_fwd: 0 = double product scan, rounded once per output (T / weights rel 5e-6 vs the
                              * reference); 1 = replay the reference's `float T_cum` recurrence sample by sample
                              * (render_utils_kernel.cu:447-451): T, weights, stop index and survivor set bit-exact */
} dvgo_scene_t;

/* Upper bound of samples per ray: ceil((far-near)/stepdist), at least 1 (t is clamped to [near,far],
 * render_utils_kernel.cu:32-33,47); ndc: ndc_samples.  Host-side helper for sizing the workspace. */
int dvgo_fused_max_steps(const dvgo_scene_t* scene);

/* ray_setup: t_min [N], n_steps [N] (int32), ray_off [N+1] (int32 exclusive scan; ray_off[N] = M0).
 * Replaces infer_t_minmax + infer_n_samples + cumsum + the .item() sync of sample_pts_on_rays
 * (render_utils_kernel.cu:190-213). */
int dvgo_fused_ray_setup(const float* rays_o, const float* rays_d, const dvgo_scene_t* scene,
                         int n_rays, float* t_min, int32_t* n_steps, int32_t* ray_off,
                         dvgo_stream_t stream);

/* march_fwd.  Slot arrays [slot_cap], survivor arrays [surv_cap] (feat [surv_cap, C]); per ray
 * alphainv_last [N]; counters[0] = number of survivors M4 (must be zeroed by the caller, e.g. with
 * dvgo_fused_zero), counters[1] = overflow flag (set if a capacity was too small).
 * k0_cl == NULL skips the k0 gather (k0_gather / k0_gather_tiles follow).  The four slot arrays are the record
 * march_bwd reads; pass all four as NULL for a forward-only (rendering) call: nothing is written per slot. */
int dvgo_fused_march_fwd(const float* rays_o, const float* rays_d, const dvgo_scene_t* scene,
                         const float* density, const float* k0_cl, int n_rays, const float* t_min,
                         const int32_t* n_steps, const int32_t* ray_off, int64_t slot_cap,
                         int64_t surv_cap, float* slot_alpha, float* slot_T, float* slot_expd,
                         int32_t* slot_code, float* feat, int32_t* s_ray, int32_t* s_slot,
                         float* s_weight, float* alphainv_last, int32_t* counters,
                         float* s_pos, dvgo_stream_t stream);

/* k0 gather / scatter over the survivor stream (one thread per survivor and 16-byte channel group).
 * march_fwd with k0_cl == NULL followed by k0_gather is equivalent to march_fwd with k0_cl; likewise
 * march_bwd with grad_k0_cl == NULL followed by k0_scatter.  Splitting them takes the k0 traffic (6x the
 * density traffic) out of the per-ray scan loops, where it is latency-bound. */
int dvgo_fused_k0_gather(const float* rays_o, const float* rays_d, const dvgo_scene_t* scene,
                         const float* k0_cl, const float* t_min, const int32_t* ray_off,
                         const int32_t* s_ray, const int32_t* s_slot, const int32_t* counters,
                         int64_t surv_cap, float* feat, dvgo_stream_t stream);
/* k0_gather whose output is the rgbnet's X~ tiles (fp16, operand layout; see dvgo_mlp_fwd below) instead of the
 * fp32 feature stream: features, the embedding share of the survivor's ray (pe_rows16 [n_rays][K1] halves as
 * dvgo_view_embedding writes them; pe_stride = the fp32 table's row length) and the zero padding, rows up to the next
 * multiple of 256 zeroed.  Same bytes as dvgo_fused_k0_gather followed by dvgo_mlp_pack_x, without the
 * 2 x M4 x C x 4 bytes of feature traffic in between. */
int dvgo_fused_k0_gather_tiles(const float* rays_o, const float* rays_d, const dvgo_scene_t* scene,
                               const float* k0_cl, const float* t_min, const int32_t* ray_off,
                               const int32_t* s_ray, const int32_t* s_slot, const int32_t* counters,
                               int64_t surv_cap, const float* s_pos, const void* pe_rows16, int pe_stride,
                               void* xt, dvgo_stream_t stream);
/* s_pos (optional in gather_tiles / scatter, [surv_cap,4] floats, 16-byte aligned): the per-survivor record
 * dvgo_fused_march_fwd writes when given the buffer -- the sample's continuous voxel coordinates (what ATen's
 * grid_sampler derives from the point, lib/dvgo.py:316-321) and the ray index as int bits.  With it the k0 kernels
 * skip the per-thread ray geometry (same floats, so the results are bit-identical); NULL recomputes it. */
int dvgo_fused_k0_scatter(const float* rays_o, const float* rays_d, const dvgo_scene_t* scene,
                          const float* t_min, const int32_t* ray_off, const int32_t* s_ray,
                          const int32_t* s_slot, const int32_t* counters, int64_t surv_cap,
                          const float* s_pos, const float* d_feat, float* grad_k0_cl, dvgo_stream_t stream);

/* rgb = sigmoid(feat) for models without rgbnet (lib/dvgo.py:512-514); rgb [M4,3], C must be 3. */
int dvgo_fused_rgb_direct(const float* feat, const int32_t* counters, int64_t surv_cap, float* rgb,
                          dvgo_stream_t stream);
/* backward of the above: d_feat = d_rgb * rgb * (1 - rgb). */
int dvgo_fused_rgb_direct_bwd(const float* rgb, const float* d_rgb, const int32_t* counters,
                              int64_t surv_cap, float* d_feat, dvgo_stream_t stream);

/* composite: rgb_acc[ray] += w*rgb (3), depth_acc[ray] += w*step (lib/dvgo.py:554-558,571-575).
 * rgb_acc [N,3], depth_acc [N] must be zeroed by the caller; depth_acc may be NULL. */
int dvgo_fused_composite(const float* rgb, const float* s_weight, const int32_t* s_ray,
                         const int32_t* s_slot, const int32_t* ray_off, const int32_t* counters,
                         int64_t surv_cap, float* rgb_acc, float* depth_acc, dvgo_stream_t stream);

/* ray_finish: rgb_marched = rgb_acc + alphainv_last*bg (lib/dvgo.py:559) in place in rgb_acc.
 * With target != NULL also the training loss pieces of run.py:377-382:
 *   G [N,3]       = dL/d rgb_marched = weight_main * 2 (rgb_marched - target) / (3 n_global)
 *   g_last [N]    = dL/d alphainv_last = bg * sum_c G + weight_entropy_last * dEntropy/dp / n_global
 *   loss_acc[0]  += weight_main * mse  + weight_entropy_last * entropy   (this rank's share)
 * n_global = number of rays in the GLOBAL batch (ray-sharded data parallel divides by it). */
int dvgo_fused_ray_finish(float* rgb_acc, const float* alphainv_last, const float* target, float bg,
                          int n_rays, int n_global, float weight_main, float weight_entropy_last,
                          float* G, float* g_last, float* loss_acc, dvgo_stream_t stream);

/* sample_grad: per survivor i of ray r (run.py:383-386 and SURVEY.md appendix C):
 *   d_rgb[i] = w_i*G[r] + weight_rgbper * 2 w_i (rgb_i - target[r]) / n_global
 *   d_w[i]   = sum_c G[r,c]*rgb_i,c                       (weights are detached in the rgbper term)
 *   loss_acc[0] += weight_rgbper * w_i * |rgb_i - target[r]|^2 / n_global
 * dzt (optional, dvgo_mlp_dztile_bytes(surv_cap) bytes): the rgbnet backward's dZ3 tiles, grad_scale * d_rgb[i] *
 * rgb_i * (1 - rgb_i) (the gradient through the sigmoid of lib/dvgo.py:539) as saturated fp16 in the operand layout,
 * rows up to the next multiple of 256 zeroed -- the input of dvgo_mlp_bwd; d_rgb may then be NULL (not written). */
int dvgo_fused_sample_grad(const float* rgb, const float* s_weight, const int32_t* s_ray,
                           const float* G, const float* target, const int32_t* counters,
                           int64_t surv_cap, int n_global, float weight_rgbper, float* d_rgb,
                           float* d_w, float* loss_acc, void* dzt, float grad_scale, dvgo_stream_t stream);

/* march_bwd: accumulates into grad_density [X,Y,Z] and grad_k0_cl [X,Y,Z,C] (fp32 atomics).
 * d_feat [M4,C] = dL/d(k0 features), d_w [M4] = dL/d(weights), g_last [N] = dL/d(alphainv_last). */
int dvgo_fused_march_bwd(const float* rays_o, const float* rays_d, const dvgo_scene_t* scene,
                         int n_rays, const float* t_min, const int32_t* n_steps,
                         const int32_t* ray_off, const float* slot_alpha, const float* slot_T,
                         const float* slot_expd, const int32_t* slot_code, const float* d_feat,
                         const float* d_w, const float* alphainv_last, const float* g_last,
                         float* grad_density, float* grad_k0_cl, dvgo_stream_t stream);

/* sweep: for every element  g = grad; if (tv && (tv_dense || g != 0)) g += TV(param_in);
 *        if (!masked || g != 0) Adam(param, g, m, v);  grad = 0.
 * `param_in` and `param_out` may alias ONLY when tv == 0 (TV reads neighbours of the old values, so
 * with TV the new parameters go to a second buffer and the caller swaps -- no extra HBM traffic).
 * layout: channels = C innermost for k0 (pass C), 1 for density.  wy/wz as in total_variation_add_grad
 * (the caller passes the un-divided weights; /6 and the wx->wz quirk are applied inside).
 * perlr (may be NULL): per-element learning-rate scale (adam_upd_with_perlr).
 * [x_begin, x_end): the x-slab this call updates (x_end < 0 = X).  All pointers address the FULL grids;
 * ray-sharded training reduce-scatters the gradient, sweeps 1/n of the grid per rank and all-gathers the
 * parameters (TV neighbours across the slab boundary come from the replicated old parameters). */
int dvgo_fused_sweep(const float* param_in, float* param_out, float* grad, float* exp_avg,
                     float* exp_avg_sq, const float* perlr, int X, int Y, int Z, int C, int x_begin,
                     int x_end, int tv, int tv_dense, float wx, float wy, float wz, int masked, int step,
                     float beta1, float beta2, float lr, float eps, dvgo_stream_t stream);

/* The same sweep as ONE kernel with the gradient exchange of ray-sharded data parallel training folded in
 * (reduce-scatter + sweep + all-gather over NVLink peer memory): for the x-slab [x_begin, x_end) this rank owns,
 * the gradient is the sum over r < n_peers of grad_peers_host[r][e] (peer loads, rank order), and the new
 * parameters are stored to param_out_peers_host[r][e] for every r (peer stores).  The two host arrays hold
 * n_peers DEVICE pointers, all mapped in this process (torch symmetric memory / CUDA IPC), index self_rank being
 * the local buffers; exp_avg / exp_avg_sq / perlr are local.  The caller orders the kernel after every rank's
 * backward pass and must not touch the gradient / output buffers again before every rank's sweep has finished.
 * param_out_multicast / grad_multicast: NVLS multicast addresses of the same buffers (one multimem.st reaches every
 * rank, one multimem.ld_reduce returns the switch-summed gradient), or NULL to use the per-peer loads / stores. */
int dvgo_fused_sweep_peer(const float* param_in, float* const* param_out_peers_host,
                          float* const* grad_peers_host, float* param_out_multicast,
                          const float* grad_multicast, int n_peers, int self_rank, float* exp_avg,
                          float* exp_avg_sq, const float* perlr, int X, int Y, int Z, int C, int x_begin,
                          int x_end, int tv, int tv_dense, float wx, float wy, float wz, int masked, int step,
                          float beta1, float beta2, float lr, float eps, dvgo_stream_t stream);

/* Layout converters at the state_dict boundary: [C,X,Y,Z] <-> [X,Y,Z,C]. */
int dvgo_grid_ncdhw_to_cl(const float* src, float* dst, int C, int64_t G, dvgo_stream_t stream);
int dvgo_grid_cl_to_ncdhw(const float* src, float* dst, int C, int64_t G, dvgo_stream_t stream);

/* rgbnet on the tensor cores (fused_mlp.cu): x = [feat (C) | pe[s_ray] (P)] -> Linear(C+P, 128) -> ReLU
 * -> Linear(128,128) -> ReLU -> Linear(128,3) -> sigmoid   (lib/dvgo.py:123-131, :524-539 with
 * rgbnet_direct=True, rgbnet_depth=3, rgbnet_width=128 -- the configs' default).
 * counters[0] = M4 (survivors), counters[1] |= 2 on a non-finite value.
 *
 * Per-survivor inputs are fp16 TILES in the tensor core's canonical shared-memory operand layout, so that a CTA
 * fetches a tile with one bulk copy (cp.async.bulk) and hands it to the MMA untouched:
 *   X~ tiles  (xt):  tile t = survivors [128 t, 128 t + 128) as [128][K1] halves, K1 = round_up(C + pe_stride, 16):
 *                    columns [0,C) k0 features, [C, C+pe_stride) the survivor's ray's row of the padded view-embedding
 *                    table pe [n_rays, pe_stride] (P embedding values, then the constant 1 that carries b1, then 0), rest 0.
 *   dZ3 tiles (dzt): [128][16] halves, columns 0..2 = grad_scale * d_rgb * rgb * (1 - rgb) (the gradient at the output
 *                    layer's pre-activation; lib/dvgo.py:539), rest 0.
 * byte offset of (row r, column c) in a tile of `cols` columns: (r/8)*(cols/8)*128 + (c/8)*128 + (r%8)*16 + (c%8)*2.
 * Rows past M4 up to the next multiple of 256 must be ZERO (the backward kernel works on tile pairs); buffers hold
 * dvgo_mlp_xtile_bytes / dvgo_mlp_dztile_bytes bytes (one tile more than ceil(surv_cap / 128)), 16-byte aligned, and
 * must be ZERO-INITIALISED by the caller once: the producers never write a 16-byte chunk that holds padding columns
 * only (X~ columns >= round_up(C + pe_stride, 8), dZ3 columns 8..15).
 * Producers: dvgo_fused_k0_gather_tiles and dvgo_fused_sample_grad write them in the fused step; dvgo_mlp_pack_x /
 * dvgo_mlp_pack_dz build them from fp32 streams (feat [surv_cap,C], s_ray [surv_cap]; rgb / d_rgb [surv_cap,3]).
 * Conversion to fp16 saturates (cvt.rn.satfinite).
 *
 * Weights: fp32 masters in torch nn.Linear layout ([out][in]) are converted ONCE per step by dvgo_mlp_pack_weights
 * into `wpack` (dvgo_mlp_wpack_bytes bytes of device memory: fp16 operand tiles, saturating conversion), which every
 * CTA of the forward / backward kernels copies.  GEMM operands are FP16, accumulation is fp32.  rgb [surv_cap,3].
 * width must be 128. */
int64_t dvgo_mlp_wpack_bytes(int C, int pe_stride);
int dvgo_mlp_pack_weights(int C, int P, int pe_stride, const float* W1, const float* b1, const float* W2,
                          const float* b2, const float* W3, const float* b3, int width, void* wpack,
                          dvgo_stream_t stream);
int64_t dvgo_mlp_xtile_bytes(int64_t surv_cap, int C, int pe_stride);
int64_t dvgo_mlp_dztile_bytes(int64_t surv_cap);
int dvgo_mlp_pack_x(const float* feat, int C, const int32_t* s_ray, const float* pe, int P, int pe_stride,
                    const int32_t* counters, int64_t surv_cap, void* xt, dvgo_stream_t stream);
int dvgo_mlp_pack_dz(const float* rgb, const float* d_rgb, float grad_scale,
                     const int32_t* counters, int64_t surv_cap, void* dzt, dvgo_stream_t stream);
int dvgo_mlp_fwd(const void* xt, int C, int P, int pe_stride, int32_t* counters, int64_t surv_cap, const void* wpack,
                 float* rgb, dvgo_stream_t stream);
/* k0 gather + rgbnet forward in ONE kernel: four producer warps per CTA build each X~ tile (trilinear k0 features of
 * lib/dvgo.py:509 + the ray's embedding row) straight into the shared-memory buffer the layer-1 MMA reads, in the
 * shadow of the tensor-core chain of the previous tiles (replaces dvgo_fused_k0_gather_tiles + dvgo_mlp_fwd, i.e.
 * F.grid_sample at lib/dvgo.py:321 + rgbnet at :536-539).  s_pos [surv_cap,4]: march_fwd's per-survivor record.
 * xt_out: NULL (rendering: the tiles never leave shared memory) or the X~ tile buffer of dvgo_mlp_xtile_bytes bytes
 * (training: written by one bulk copy per tile, covering whole tile pairs, for dvgo_mlp_bwd). */
int dvgo_mlp_fwd_gather(const dvgo_scene_t* scene, const float* k0_cl, const float* s_pos, const void* pe_rows16,
                        int P, int pe_stride, int32_t* counters, int64_t surv_cap, const void* wpack, float* rgb,
                        void* xt_out, dvgo_stream_t stream);
/* Same as dvgo_mlp_fwd; if `timeline` is non-NULL, CTA 0 records clock64() at every phase boundary of its
 * first tiles into timeline[0..63] (thread 0) and timeline[64..127] (thread 255) -- kernel-author tooling. */
int dvgo_mlp_fwd_timed(const void* xt, int C, int P, int pe_stride, int32_t* counters, int64_t surv_cap,
                       const void* wpack, float* rgb, long long* timeline, dvgo_stream_t stream);
/* Backward with forward recompute: d_feat [surv_cap,C] = dL/dfeat, and gW*, gb* += weight gradients
 * (accumulated in TMEM per CTA, flushed with atomics; the caller zeroes them).  grad_scale: the power of two the dZ3
 * tiles were scaled by (keeps the FP16 backward operands in normal range); removed exactly in the fp32 epilogues.
 * `wpack` and `xt` must be the ones the forward used. */
int dvgo_mlp_bwd(const void* xt, const void* dzt, int C, int P, int pe_stride, int32_t* counters, int64_t surv_cap,
                 const void* wpack, float grad_scale, float* d_feat, float* gW1, float* gb1, float* gW2, float* gb2,
                 float* gW3, float* gb3, dvgo_stream_t stream);
/* dvgo_mlp_bwd with an optional in-kernel timeline (CTA 0: epilogue thread 0 -> timeline[0..63], issuer ->
 * timeline[64..127], clock64 at every phase boundary of the first tiles) -- kernel-author tooling. */
int dvgo_mlp_bwd_timed(const void* xt, const void* dzt, int C, int P, int pe_stride, int32_t* counters,
                       int64_t surv_cap, const void* wpack, float grad_scale, float* d_feat, float* gW1, float* gb1,
                       float* gW2, float* gb2, float* gW3, float* gb3, long long* timeline, dvgo_stream_t stream);

/* Cross-GPU barrier on the stream without a collective library: flags_peers_host[r] = rank r's array of n_peers int32
 * flags (symmetric memory, zero-initialised, peer-mapped); the kernel stores `epoch` into slot self_rank of every rank's
 * array (release.sys) and waits until every slot of the local array has reached it (acquire.sys).  epoch must grow by
 * one per call on every rank.  A rank that never arrives makes the others trap after ~4 s instead of hanging. */
int dvgo_peer_barrier(int32_t* const* flags_peers_host, int n_peers, int self_rank, int epoch, dvgo_stream_t stream);
/* Adam (lib/masked_adam.py:60-71, dense variant) on a small replicated tensor whose gradient is the sum over ranks,
 * read from the peers' buffers in rank order (grad_multicast non-NULL: one multimem.ld_reduce per element instead).
 * The rgbnet's 22 K parameters in ray-sharded training: every rank computes the identical update. */
int dvgo_adam_upd_peer(float* param, const float* const* grad_peers_host, const float* grad_multicast, int n_peers,
                       float* exp_avg, float* exp_avg_sq, int64_t N, int step, float beta1, float beta2, float lr,
                       float eps, dvgo_stream_t stream);

/* Tensor-core self test (one CTA): D[128,N] = A * B^T with tcgen05.mma kind::f16 (fp16 operands), for each operand
 * orientation the rgbnet kernels use.  a_mn=0: A is [128][K]; a_mn=1: A is [K][128]; b_mn=0: B is
 * [N][K]; b_mn=1: B is [K][N].  N % 16 == 0, N <= 256, K % 16 == 0, K <= 128.  D is [128][N]. */
int dvgo_tc_selftest(const float* A, const float* B, float* D, int N, int K, int a_mn, int b_mn,
                     dvgo_stream_t stream);

/* Descriptor probe (debug aid for the kernel author): A [128][K] K-major, the B operand region is
 * filled verbatim from Braw [nwords] and described with the given LBO/SBO/k-step (bytes). */
/* View-direction encoding of lib/dvgo.py:524-525 written into the padded table the rgbnet kernels read:
 * out [n_rays, stride] = [viewdirs | sin(v*f) | cos(v*f) | 1 | 0...], stride >= 3 + 6*n_freq + 1.
 * rows16 (optional, [n_rays][K1] halves, K1 = round_up(C + stride, 16), ZERO-initialised by the caller): the same
 * values as saturated fp16 in columns [C, C + stride) -- the ray's share of an X~ row, which
 * dvgo_fused_k0_gather_tiles copies 16 bytes at a time instead of converting it again for every survivor. */
int dvgo_view_embedding(const float* viewdirs, const float* freq, int n_freq, int64_t n_rays, int stride,
                        float* out, void* rows16, int C, dvgo_stream_t stream);

/* Tensor-core issue-rate probe (tools/mma_rate.py): `reps` x `ksteps` M=128 MMAs of width N issued by one thread per
 * CTA from zero-filled shared memory with the given descriptor fields; out[cta] = cycles. */
int dvgo_tc_rate(int ctas, int N, int ksteps, int reps, int a_mn, int b_mn, int a_lbo, int a_sbo, int a_kstep,
                 int b_lbo, int b_sbo, int b_kstep, int layout, int n_accum, long long* out, dvgo_stream_t stream);
int dvgo_tc_probe(const float* A, const float* Braw, float* D, int N, int K, int b_mn, int lbo, int sbo,
                  int kstep, int nwords, dvgo_stream_t stream);

/* Tensor-memory A-operand probe (tools/ts_probe.py): D[128,N] = A[128,K] * B^T with A staged in TMEM by tcgen05.st
 * (fp16 pairs packed along K) and B in shared memory ([N][K], or [K][N] when b_mn); cycles[0] / cycles[1] = clock64
 * cycles of `reps` x K/16 MMAs in the TMEM-A form and in the shared-memory-A form. */
int dvgo_tc_ts_probe(const float* A, const float* B, float* D, int N, int K, int b_mn, int reps, long long* cycles,
                     dvgo_stream_t stream);

/* Tensor pipe vs SIMT memory-path contention probe (tools/contention.py): out[0] = SIMT cycles, out[1] = MMA cycles;
 * gbuf: >= 1 MiB of device memory. */
int dvgo_tc_contention(int mma_mode, int n_mma, int simt_mode, int reps, const void* gbuf, long long* out,
                       dvgo_stream_t stream);

/* TMEM -> register read-rate probe (tools/ldtm_rate.py): out[0] = cycles, out[1] = bytes for nwarps x reps reads. */
int dvgo_tc_ldtm_rate(int nwarps, int reps, int mode, long long* out, dvgo_stream_t stream);

/* Zero `n` 4-byte words (counters, accumulators) on the stream. */
int dvgo_fused_zero(void* ptr, int64_t n_words, dvgo_stream_t stream);

/* Start of a fused call: zblock = counters[2] | accumulators (n_words 4-byte words, n_words >= 2).  Before zeroing,
 * the previous call's counters are folded into stats (4 x int64, device; may be NULL): [0] += survivors,
 * [1] += 1, [2] |= overflow flag, [3] = max(survivors).  Lets the host learn the data-dependent survivor count of
 * the timed steps (the MLP work term of the roofline) and a capacity overflow without synchronising per step. */
int dvgo_fused_step_begin(void* zblock, int64_t n_words, long long* stats, dvgo_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* DVGO_B200_FUSED_H_ */
