/*
 * dvgo_b200_prep.h -- C ABI of the "next" rows of the hot-path scope table:
 * the callers either side of the per-iteration path, built on the same sampling / occupancy /
 * trilinear device code as the fused kernels.
 *
 *   N2  ray generation                 lib/ray_utils.py:9-47 (get_rays), :62-85 (ndc_rays, get_rays_of_a_view)
 *   N1  training-ray preparation       lib/ray_utils.py:146-183 (get_training_rays_in_maskcache_sampling),
 *                                      lib/dvgo.py:412-423 (hit_coarse_geo)
 *   N3  whole-grid / all-ray sweeps    lib/dvgo.py:265-295 (voxel_count_views), run.py:330-332 (occupancy
 *                                      refresh), lib/dvgo.py:229-263 (scale_volume_grid)
 *
 * Conventions as in dvgo_b200.h: device pointers unless the name ends in _host, explicit stream,
 * 0 / cudaError_t / DVGO_EINVAL return, no allocation, no synchronisation.
 */
#ifndef DVGO_B200_PREP_H_
#define DVGO_B200_PREP_H_

#include <stdint.h>

#include "dvgo_b200.h"
#include "dvgo_b200_fused.h"

#ifdef __cplusplus
extern "C" {
#endif

/* One pinhole view (host struct, passed by value to the kernels). */
typedef struct dvgo_view {
  int H, W;
  float fx, fy, cx, cy;   /* K[0][0], K[1][1], K[0][2], K[1][2] as fp32 (torch rounds the numpy scalars to the
                             tensor dtype, lib/ray_utils.py:32-34) */
  float c2w[12];          /* camera-to-world rows 0..2, row-major [3][4] (c2w[:3,:4]) */
  int inverse_y, flip_x, flip_y;
  int mode;               /* 0 = 'lefttop', 1 = 'center' (pixel centre +0.5); 'random' is host-side only */
  int ndc;                /* apply ndc_rays(H, W, focal=K[0][0], near=1.) */
  float ndc_sx, ndc_sy;   /* -1/(W/(2 focal)), -1/(H/(2 focal)): evaluated by the host in Python arithmetic
                             exactly as lib/ray_utils.py:68-69,73-74 does, then rounded to fp32 */
} dvgo_view_t;

/* N2  get_rays_of_a_view for pixels [pix_begin, pix_begin + n_pix) of the row-major H x W image:
 * rays_o, rays_d, viewdirs [n_pix,3].  viewdirs is computed from the pre-NDC direction
 * (lib/ray_utils.py:82-84).  Any output pointer may be NULL. */
int dvgo_rays_of_view(const dvgo_view_t* view_host, int64_t pix_begin, int64_t n_pix, float* rays_o,
                      float* rays_d, float* viewdirs, dvgo_stream_t stream);

/* N1  hit_coarse_geo (lib/dvgo.py:412-423): hit[r] = any sample of ray r that is inside the bbox AND in
 * occupied space (scene->mask).  Uses scene->{xyz_min,xyz_max,mask*,near,far,stepdist}.  Bit-exact class. */
int dvgo_hit_coarse_geo(const dvgo_scene_t* scene, const float* rays_o, const float* rays_d, int64_t n_rays,
                        uint8_t* hit, dvgo_stream_t stream);

/* N1 fused with N2: the same test for every pixel of a view, rays generated on the fly (no [H*W,3] tensors).
 * hit [H*W]. */
int dvgo_view_hit_coarse_geo(const dvgo_view_t* view_host, const dvgo_scene_t* scene, uint8_t* hit,
                             dvgo_stream_t stream);

/* N1  the compaction of get_training_rays_in_maskcache_sampling (lib/ray_utils.py:166-173): for every pixel p
 * with hit[p], row q = *top + pos_incl[p] - 1 of the output buffers receives img[p], rays_o, rays_d, viewdirs of
 * that pixel (rays regenerated on the fly).  pos_incl = inclusive prefix sum of hit (int64, [H*W]); top = device
 * scalar (rows already used by earlier views), so a whole training set is prepared without a host sync.
 * img may be NULL (then rgb_tr is not written). */
int dvgo_view_gather_rays(const dvgo_view_t* view_host, const uint8_t* hit, const int64_t* pos_incl,
                          const int64_t* top, const float* img, float* rgb_tr, float* rays_o_tr,
                          float* rays_d_tr, float* viewdirs_tr, dvgo_stream_t stream);

/* N3  voxel_count_views inner loop (lib/dvgo.py:276-291): for every ray and every i < n_samples the point
 * o + d * (t_min + (stepdist * i) / |d|) scatters its 8 trilinear weights into acc [X,Y,Z] (the backward of
 * grid_sample on a grid of ones, zero padding).  t_min = clamp(slab entry, near, far). */
int dvgo_voxel_count_scatter(const float* rays_o, const float* rays_d, int64_t n_rays, const float* xyz_min,
                             const float* xyz_max, int X, int Y, int Z, float near, float far, float stepdist,
                             int n_samples, float* acc, dvgo_stream_t stream);
/* N3  lib/dvgo.py:292-293 per view: count += (acc > 1); acc = 0. */
int dvgo_voxel_count_commit(float* acc, float* count, int64_t n, dvgo_stream_t stream);

/* N3  occupancy refresh (run.py:330-332; also lib/dvgo.py:254-259): mask_out = mask_in AND
 * (maxpool3x3x3(raw2alpha(density, act_shift, interval)) > thres); mask_in may be NULL (treated as all true),
 * mask_out may alias mask_in.  alpha_tmp [X*Y*Z] is scratch. */
int dvgo_alpha_maxpool_mask(const float* density, int X, int Y, int Z, float act_shift, float interval,
                            float thres, const uint8_t* mask_in, uint8_t* mask_out, float* alpha_tmp,
                            dvgo_stream_t stream);

/* N3  F.interpolate(grid, size, mode='trilinear', align_corners=True) of lib/dvgo.py:236-241:
 * src [C,X,Y,Z] -> dst [C,X2,Y2,Z2] (ATen upsample_trilinear3d arithmetic). */
int dvgo_resize_trilinear(const float* src, int C, int X, int Y, int Z, float* dst, int X2, int Y2, int Z2,
                          dvgo_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* DVGO_B200_PREP_H_ */
