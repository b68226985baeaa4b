// kernels.h — host-callable launchers of the CUDA kernels (one per translation unit below).
#pragma once
#include <vector>
#include <cuda_runtime.h>
#include <stdint.h>

#include "slb_dev.h"

namespace slbk {

// k_render.cu
void launch_setup(const DView* views, const DFrame* frames, const DBinDraw* bdraws, const uint32_t* chunk_draw, uint32_t n_chunks,
                  uint32_t* tile_count, PairRec* survivors, uint32_t* counters, uint32_t normal_cap, uint32_t huge_cap, int direct_max,
                  int warp_max, cudaStream_t s);
void launch_emit(const DView* views, const PairRec* survivors, uint32_t n_survivors, const PairRec* huge, uint32_t n_huge,
                 uint32_t* tile_count, PairRec* pairs, uint32_t capacity, cudaStream_t s);
void launch_scan(uint32_t* count, uint32_t* off, ActiveTile* active, unsigned long long* block_sums, uint32_t* totals,
                 const uint32_t* counters, uint32_t n, cudaStream_t s);
struct RasterGrid { uint32_t n_active, n_cam_tiles, tiles_per_cam, n_cam_views, tiles_per_shadow; };
void launch_raster(bool frag_test, const DView* views, const DFrame* frames, const DDraw* draws, const ActiveTile* active,
                   const PairRec* pairs, RasterGrid g, cudaStream_t s);
void launch_shade(const DFrame* frames, const DDraw* draws, int n_frames, int W, int H, bool lean, cudaStream_t s);

// k_post.cu
void upload_ssao_tables(const float* noise16x3, const float* kernel64x3);
void launch_background(const DFrame* frames, int n_frames, int W, int H, cudaStream_t s);
void launch_downsample(const float4* src, int sw, int sh, float4* dst, int n_frames, size_t src_stride, size_t dst_stride, cudaStream_t s);
void launch_ssao(const DFrame* frames, int n_frames, int W, int H, cudaStream_t s);
void launch_ssao_apply_tonemap(const DFrame* frames, int n_frames, int W, int H, cudaStream_t s);

// k_assets.cu
void launch_repack_vertices(const uint8_t* verts68, uint32_t n, float4* pos4, float4* attr, float4* col4, cudaStream_t s);
void launch_unpack_vertices(const float4* pos4, const float4* attr, const float4* col4, uint32_t n, uint8_t* verts68, cudaStream_t s);
void launch_index_max(const uint32_t* idx, uint32_t n, uint32_t* out, cudaStream_t s);
void launch_vertex_delta(const int32_t* ids, uint32_t n, const float* dpos, const float* dcol, float4* pos4, float4* col4, uint32_t n_vertices,
                         uint32_t* err, cudaStream_t s);
void launch_set_positions(const float* p3, uint32_t n, float4* pos4, cudaStream_t s);
void launch_set_colors(const float* c4, uint32_t n, float4* col4, cudaStream_t s);
void launch_recompute_normals(const float4* pos4, const uint32_t* idx, uint32_t n_faces, const uint32_t* adj_off, const uint32_t* adj_face,
                              float4* face_n, uint32_t n_vertices, float4* attr, cudaStream_t s);
void launch_expand_rgba(const uint8_t* src, int channels, uint8_t* dst, size_t n_texels, cudaStream_t s);
void launch_mip_level(const uint8_t* src, int sw, int sh, uint8_t* dst, int dw, int dh, cudaStream_t s);
void launch_cube_mip(const float4* src, int ssize, float4* dst, cudaStream_t s);
void launch_equirect_to_cube(const float* equirect, int ew, int eh, float4* cube, int size, cudaStream_t s);
void launch_irradiance(const DLightMap* lm, float4* out, int size, float lod, cudaStream_t s);
void launch_prefilter(const DLightMap* lm, float4* out, int size, float roughness, int n_samples, float env_resolution, cudaStream_t s);
void launch_brdf_lut(float4* out, int size, int n_samples, cudaStream_t s);

// k_diff.cu
void launch_sobel_valid_mask(const int16_t* inst, const float* depth, uint8_t* valid, int H, int W, cudaStream_t s);
void launch_dilate_object_mask(const uint8_t* mask, const uint8_t* valid, const float* coords, int coord_stride, uint8_t* mask_out,
                               float* coords_out, int H, int W, cudaStream_t s);

size_t pose_grad_partial_floats(int n_obj, int H, int W);
void launch_pose_grad(const uint8_t* rgb, const int16_t* inst, const float* coord, const float* grad_img, const float* params, int n_obj,
                      float* partial, float* out, int H, int W, cudaStream_t s);

// k_camera.cu
void launch_camera_stage1(const float* in_f, const uint8_t* in_u8, float* out, const slb_camera_params* params, int n, int H, int W,
                          cudaStream_t s);
void launch_camera_stage2(const float* in, float* out, float sigma, int n, int H, int W, cudaStream_t s);

// k_png.cu
void png_upload_tables();
int png_segments(int W, int channels, int bpc);
size_t png_seg_bound();
size_t png_file_bound(int H, int W, int channels, int bpc);
size_t png_row_info_bytes();
void launch_png_encode(const uint8_t* images, int n, int H, int W, int channels, int bpc, uint8_t* rows_scratch, void* row_info,
                       uint32_t* row_offset, uint8_t* out, size_t out_stride, uint32_t* sizes, cudaStream_t s);


// k_jpeg.cu
size_t jpeg_stream_words(int H, int W, int channels);
size_t jpeg_coef_bytes(int H, int W, int channels);
size_t jpeg_blocks(int H, int W, int channels);
size_t jpeg_file_bound(int H, int W, int channels);
struct JpegQuant { uint16_t div[2][64]; };   // quantisation divisors (8 * Q, natural order) of one quality: a kernel parameter
std::vector<uint8_t> jpeg_prepare(int H, int W, int channels, int quality, JpegQuant* quant, cudaStream_t s);
void launch_jpeg_encode(const uint8_t* images, int n, int H, int W, int channels, const JpegQuant& quant, int16_t* coefs, uint32_t* bit_off, uint32_t* total_bits,
                        uint32_t* stream, const uint8_t* header_dev, int header_len, uint8_t* out, size_t out_stride, uint32_t* sizes,
                        cudaStream_t s);

}  // namespace slbk
