/*
 * agile3d_b200 — C-ABI of the Blackwell (sm_100a) hot path of AGILE3D.
 *
 * The reference (ywyue/AGILE3D) is pure Python: it has no FFI of its own.  What it binds for this path is
 * the third-party MinkowskiEngine extension (coordinate manager, kernel maps, sparse convolution) plus ATen
 * ops (nn.MultiheadAttention, LayerNorm, matmul).  Each entry point below names the reference call site /
 * MinkowskiEngine operator it replaces.  INTEGRATION.md shows the ctypes stub a maintainer would add.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer owned by the caller (torch's allocator), 16-byte aligned;
 *   - every call is asynchronous on `stream` (a cudaStream_t passed as void*) and allocates nothing;
 *     scratch memory comes from a caller-supplied workspace whose size the *_workspace_bytes() query gives;
 *   - return value: 0 = ok, negative = error (AG3D_E_*), message via ag3d_last_error() (thread-local);
 *   - row-major fp32 features [N, C] with an explicit leading dimension (`*_ld`, in floats) so that a layer can
 *     read or write a channel slice of a wider buffer (this is how `me.cat` costs nothing);
 *   - voxel coordinates are int32 [N,4] = (batch, x, y, z), |x|,|y|,|z| < 32768, 0 <= batch < 65535;
 *   - neighbour tables ("kernel maps") are int32 [K, N_out], offset-major: nbr[k*N_out + o] is the input row
 *     that output row o sees through kernel offset k, or -1.  Offset order is MinkowskiEngine's (x fastest,
 *     odd kernels centred, even kernels one-sided).
 */
#ifndef AGILE3D_B200_H_
#define AGILE3D_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define AG3D_ABI_VERSION 7

#define AG3D_OK 0
#define AG3D_E_INVALID (-1)   /* bad argument (shape, alignment, unsupported size) */
#define AG3D_E_CUDA (-2)      /* a CUDA runtime call failed */
#define AG3D_E_WORKSPACE (-3) /* workspace too small */

/* spconv / stem `flags` */
#define AG3D_RELU 1
/* feature format of the input / output / residual rows: default fp32; *_SPLIT = every 32-channel slab stored as
 * 64 B of bf16 hi + 64 B of bf16 lo (x = hi + lo, relative error <= 2^-17), same bytes and leading dimension as
 * fp32.  The tensor-core path gathers split rows with cp.async and no conversion work; the fp32 FFMA path takes
 * fp32 rows only.                                                                                              */
#define AG3D_IN_SPLIT 2
#define AG3D_OUT_SPLIT 4
#define AG3D_RES_SPLIT 8
/* spconv `algo` */
#define AG3D_ALGO_AUTO 0
#define AG3D_ALGO_SIMT 1 /* exact fp32 FFMA implicit GEMM */
#define AG3D_ALGO_TC 2   /* tcgen05 bf16x3 (hi*hi + hi*lo + lo*hi, fp32 accumulate in TMEM) implicit GEMM */
#define AG3D_ALGO_TC_PACKED 3 /* same arithmetic, MMA rows = the present (input,output) pairs of a 256-row super tile
                                 (csrc/spconv_pk.cu); split input rows, K >= 2, cout in {32,64,96,128}.  AUTO picks
                                 it for levels with enough rows, AG3D_ALGO_TC always means the dense-tile kernel */

typedef void* ag3d_stream_t; /* cudaStream_t */

int ag3d_abi_version(void);
const char* ag3d_last_error(void);
/* sm count / compute capability of the current device */
int ag3d_device_info(int32_t* sm_count, int32_t* cc_major, int32_t* cc_minor);
/* number of CUDA kernels this library has launched in this process so far (bench.py's `gpu_launches`) */
int64_t ag3d_kernel_launches(void);

/* ---- coordinate hashing -------------------------------------------------------------------------------
 * Replaces MinkowskiEngine CoordinateMapGPU::insert_and_map reached from ME.SparseTensor(coordinates=...)
 * (reference engine.py:47-51, eval_multi_obj.py:94-98).  The table is open addressing with 16-byte slots
 * {uint64 key, int32 first_row, int32 row}; capacity is a power of two >= 2*n.
 * status[0] += number of duplicate rows, status[1] += number of out-of-range rows (caller zeroes it).      */
int64_t ag3d_hash_capacity(int64_t n);
int ag3d_hash_build(const int32_t* coords, int64_t n, void* table, int64_t cap, int32_t* status,
                    ag3d_stream_t stream);

/* ---- stride-2 coordinate map --------------------------------------------------------------------------
 * Replaces CoordinateMap::stride reached from every conv(kernel 2, stride 2) (models/res16unet.py:51-59,
 * 70-78,89-97,108-116) and MinkowskiAvgPooling(2,2) (models/agile3d.py:172-173).
 * out_coords = unique(floor(c / new_stride) * new_stride), rows in order of first occurrence while scanning
 * the fine rows (the canonical order).  Also builds the coarse level's hash table and
 * parent[i] = coarse row of fine row i.  *out_n (device int32) receives the number of coarse rows;
 * out_coords must have room for n rows.                                                                    */
size_t ag3d_downsample_workspace_bytes(int64_t n);
int ag3d_downsample(const int32_t* coords, int64_t n, int32_t new_stride, void* coarse_table, int64_t cap,
                    int32_t* parent, int32_t* out_coords, int32_t* out_n, void* ws, size_t ws_bytes,
                    ag3d_stream_t stream);

/* Internal row order of a coordinate level (no counterpart in the reference: MinkowskiEngine's row order is
 * unspecified).  perm[new] = old row, inv[old] = new row; rows are sorted scene by scene (coords[:,0]) by
 * the 27-bit pattern of their existing 3x3x3 neighbours (nbr [K, n]), stably.  ag3d_permute_map rewrites a neighbour table
 * into such an order: out[k][i] = inv_in[nbr[k][perm_out[i]]] (-1 stays; NULL = identity).                      */
size_t ag3d_row_order_workspace_bytes(int64_t n);
int ag3d_row_order(const int32_t* nbr, int32_t K, int64_t n, const int32_t* coords, int32_t* perm, int32_t* inv, void* ws,
                   size_t ws_bytes, ag3d_stream_t stream);
int ag3d_permute_map(const int32_t* nbr, int32_t K, int64_t n_out, const int32_t* perm_out, const int32_t* inv_in,
                     int32_t* out, ag3d_stream_t stream);
/* ag3d_downsample with the input row count on the DEVICE (n_dev; n_max bounds it and sizes grids, table and
 * workspace): lets the four coordinate levels of the U-Net be built back to back with ONE host read-back of all
 * counts instead of a synchronisation per level.                                                            */
int ag3d_downsample_dev(const int32_t* coords, int64_t n_max, const int32_t* n_dev, int32_t new_stride, void* coarse_table,
                        int64_t cap, int32_t* parent, int32_t* out_coords, int32_t* out_n, void* ws, size_t ws_bytes,
                        ag3d_stream_t stream);
/* Row ranges of the scenes of a batched coordinate list (ME.utils.batched_coordinates order):
 * offsets[b] = first row of scene b for b = 0 .. max_scenes (= n past the last scene); offsets[max_scenes + 1] =
 * number of rows out of batch order (must be 0).  Device int32[max_scenes + 2].                               */
int ag3d_scene_offsets(const int32_t* coords, int64_t n, int32_t max_scenes, int32_t* offsets, ag3d_stream_t stream);
/* ---- kernel maps --------------------------------------------------------------------------------------
 * Replaces CoordinateMapGPU::kernel_map (first conv of each (level, kernel) pair; cached afterwards).
 * Generic region: for every output coordinate o and offset k (ksize^3 offsets scaled by
 * dilation*in_tensor_stride), nbr[k][o] = row of (o + off_k) in the input table or -1.
 * pair_count (nullable, device int32[K], caller zeroes) receives the number of valid pairs per offset.      */
int ag3d_kernel_map(const int32_t* out_coords, int64_t n_out, const void* in_table, int64_t cap, int32_t ksize,
                    int32_t in_tensor_stride, int32_t dilation, int32_t* nbr, int32_t* pair_count,
                    ag3d_stream_t stream);
/* Transposed kernel-2/stride-2 map (MinkowskiConvolutionTranspose, models/modules/common.py:158-188):
 * nbr[k][f] = parent[f] if k is the offset of fine voxel f inside its parent cell, else -1.                 */
int ag3d_kernel_map_transposed(const int32_t* fine_coords, const int32_t* parent, int64_t n_fine,
                               int32_t fine_stride, int32_t* nbr, ag3d_stream_t stream);

/* ---- sparse convolution forward -----------------------------------------------------------------------
 * Replaces MinkowskiConvolution / MinkowskiConvolutionTranspose forward (ConvolutionForwardKernelGPU) and
 * the un-fused MinkowskiBatchNorm(eval) + MinkowskiReLU + residual add + me.cat around it
 * (models/modules/resnet_block.py:48-64, models/res16unet.py:222-295).
 *   out[o, :] = act( scale * sum_k in[nbr[k][o], :] @ W[k]  + shift  (+ residual[o, :]) )
 * weight is MinkowskiEngine's [K, cin, cout]; nbr == NULL means K == 1 on the identity map (1x1 conv).
 * scale/shift are the folded BatchNorm (or NULL/bias); residual is nullable.  cin, cout multiples of 32.
 * weight_tc (nullable) is the same weight pre-split into bf16 hi/lo stage images by
 * ag3d_spconv_tc_prepare_weight; AG3D_ALGO_AUTO takes the tensor-core path when it is given.                */
size_t ag3d_spconv_tc_weight_bytes(int32_t K, int32_t cin, int32_t cout);
int ag3d_spconv_tc_prepare_weight(const float* weight, int32_t K, int32_t cin, int32_t cout, void* weight_tc,
                                  ag3d_stream_t stream);
/* workspace (split-K partial sums of the tensor-core path on levels with few rows); 0 when none is needed */
size_t ag3d_spconv_workspace_bytes(int64_t n_out, int32_t K, int32_t cin, int32_t cout);
int ag3d_spconv_fwd(const float* in, int32_t in_ld, int32_t cin, const int32_t* nbr, int32_t K, int64_t n_out,
                    const float* weight, const void* weight_tc, int32_t cout, const float* scale, const float* shift,
                    const float* residual, int32_t res_ld, float* out, int32_t out_ld, int32_t flags,
                    int32_t algo, void* ws, size_t ws_bytes, ag3d_stream_t stream);
/* Same operator with the number of rows of `in` stated (n_in; with nbr == NULL it equals n_out).  Knowing the extent
 * of the input lets the tensor-core path describe it to the TMA engine (tensor map) and gather the neighbour rows
 * with tile::gather4 instead of per-thread cp.async; n_in == 0 means "unknown" and is what ag3d_spconv_fwd passes. */
int ag3d_spconv_fwd_rows(const float* in, int64_t n_in, int32_t in_ld, int32_t cin, const int32_t* nbr, int32_t K,
                         int64_t n_out, const float* weight, const void* weight_tc, int32_t cout, const float* scale,
                         const float* shift, const float* residual, int32_t res_ld, float* out, int32_t out_ld,
                         int32_t flags, int32_t algo, void* ws, size_t ws_bytes, ag3d_stream_t stream);
/* Stem: conv0p1s1 (3 -> 32 channels, kernel 5, models/res16unet.py:39-47) evaluated directly against the
 * hash table (no 125-column neighbour table is materialised) with folded bn0 + ReLU.                        */
int ag3d_stem_conv_fwd(const int32_t* coords, const float* feats, int64_t n, const void* table, int64_t cap,
                       int32_t ksize, const float* weight, const float* scale, const float* shift, float* out,
                       int32_t out_ld, int32_t flags, ag3d_stream_t stream);

/* dst[i] = src[idx[i]] for rows of row_bytes bytes (a multiple of 4): the internal row order of ag3d_row_order applied
 * to coordinates / raw xyz, and its inverse applied to the mask logits handed back to the caller.             */
int ag3d_gather_rows(const void* src, int32_t row_bytes, const int32_t* idx, int64_t n, void* dst, ag3d_stream_t stream);
/* Bricks: the 4x4x4 block of full-resolution cells under one tensor-stride-4 voxel (a level-2 row of the U-Net's
 * coordinate hierarchy, models/res16unet.py:222-256).  brick_rows[n_bricks][64] = full-resolution row of cell
 * (x&3 | (y&3)<<2 | (z&3)<<4) of the brick, -1 = empty; parent01 / parent12 are the parent_idx outputs of the two
 * ag3d_downsample steps.  The *_bricks stem variants take the tensor-stride-4 table and this array instead of the
 * full-resolution table: a 5^3 window is 8 probes + 125 short reads instead of 125 probes; results are identical. */
int ag3d_brick_rows(const int32_t* coords, const int32_t* parent01, const int32_t* parent12, int64_t n,
                    int64_t n_bricks, int32_t* brick_rows, ag3d_stream_t stream);
int ag3d_stem_conv_fwd_bricks(const int32_t* coords, const float* feats, int64_t n, const void* table2, int64_t cap2,
                              const int32_t* brick_rows, int32_t ksize, const float* weight, const float* scale,
                              const float* shift, float* out, int32_t out_ld, int32_t flags, ag3d_stream_t stream);

/* ---- fourier positional encoding ----------------------------------------------------------------------
 * Replaces Agile3d.get_pos_encs -> PositionEmbeddingCoordsSine.get_fourier_embeddings
 * (models/agile3d.py:141-161, models/position_embedding.py:13-41,123-152) for the full-resolution level:
 * per scene b (rows scene_offsets[b]..scene_offsets[b+1]): u = (xyz - min) / (max - min);
 * out = [sin(2*pi*u @ B), cos(2*pi*u @ B)]  with B = gauss_B [3, d/2].  range_out (nullable) receives
 * [n_scenes, 6] = (min xyz, max xyz) for the click-query encodings.                                         */
size_t ag3d_posenc_workspace_bytes(int32_t n_scenes);
int ag3d_fourier_posenc(const float* xyz, const int32_t* scene_offsets_host, int32_t n_scenes,
                        const float* gauss_B, int32_t d_pos, float* out, float* range_out, void* ws,
                        size_t ws_bytes, ag3d_stream_t stream);
/* ... and, additionally, the same encodings as "split" rows (out_split: [N, d_pos] as 32-channel slabs of 64 B bf16 hi |
 * 64 B bf16 lo) for the TMA-fed decoder kernels (ag3d_c2s_attn_fwd_split / ag3d_s2c_mask_fwd_split).            */
int ag3d_fourier_posenc_split(const float* xyz, const int32_t* scene_offsets_host, int32_t n_scenes,
                              const float* gauss_B, int32_t d_pos, float* out, float* out_split, float* range_out,
                              void* ws, size_t ws_bytes, ag3d_stream_t stream);

/* ---- interactive loop around forward_mask and the voxelisation front end -----------
 * ag3d_click_pred:     pred[v] = argmax_o logits[v, o] (first maximum; logits == NULL: all zeros, the first round of
 *                      eval_multi_obj.py:120-121), then pred[click_rows[i]] = click_objs[i] (eval_multi_obj.py:136-138).
 * ag3d_scene_iou:      counts[o] = (|p == o & l == o|, |p == o|, |l == o|) for p = pred[inverse_map] (NULL: identity),
 *                      l = labels_full: the ingredients of utils/seg.py:9-17,44-59 (mean_iou_scene).  uint64 [n_obj, 3].
 * ag3d_click_simulate: utils/seg.py:173-226.  Error clusters are the (gt, pred) pairs with gt != pred; the size of a
 *                      cluster is the largest distance of one of its voxels to the nearest voxel outside the cluster
 *                      (xyz: [nv, 3] fp32), its click the first voxel attaining it.  Clusters are ranked by size,
 *                      descending (ties in ascending 96 gt + 11 pred, the reference's cluster id); the first top_n
 *                      (top_n < 0: all) are taken in the order perm[0], perm[1], ... of their ranks (perm == NULL:
 *                      rank order; the reference shuffles the selected list with random.shuffle).
 *                      out int32 [1 + 4 max_new]: n, then n x (voxel row, object id = gt of the voxel, 32 gt + pred),
 *                      then (from 1 + 3 max_new) the n sizes as float bits.
 * ag3d_quantize_points: coords[i] = (batch_index, floor(points[i] / quantization_size)) as int32 [n, 4]; *status +=
 *                      points outside +-32767.  ME.utils.sparse_quantize = this + ag3d_downsample(stride 1) (unique voxels
 *                      in first-occurrence order, parent = inverse_map) + ag3d_first_rows (unique_map).
 * ag3d_first_rows:     unique_map[j] = min{ i : parent[i] == j }  (int64 [m]).                                        */
int ag3d_click_pred(const float* logits, int32_t n_obj, int64_t nv, const int32_t* click_rows, const int32_t* click_objs,
                    int32_t n_clicks, int32_t* pred, ag3d_stream_t stream);
int ag3d_scene_iou(const int32_t* pred, const int64_t* inverse_map, const int32_t* labels_full, int64_t n_full, int32_t n_obj,
                   uint64_t* counts, ag3d_stream_t stream);
size_t ag3d_click_simulate_workspace_bytes(int64_t nv);
int ag3d_click_simulate(const int32_t* pred, const int32_t* gt, const float* xyz, int64_t nv, int32_t top_n,
                        const int32_t* perm, int32_t max_new, int32_t* out, void* ws, size_t ws_bytes, ag3d_stream_t stream);
int ag3d_quantize_points(const float* points, int64_t n, float quantization_size, int32_t batch_index, int32_t* coords,
                         int32_t* status, ag3d_stream_t stream);
int ag3d_first_rows(const int32_t* parent, int64_t n, int64_t m, int64_t* unique_map, ag3d_stream_t stream);

/* ---- click-query side of a decoder layer (K11) -------------------------------------------------------------
 * Replaces the O(Nq) torch calls between the two voxel-streaming kernels of a layer (models/agile3d.py:273-325):
 * the q/k/v/out projections of CrossAttentionLayer / SelfAttentionLayer (models/modules/attention_block.py:28-38,
 * 86-98), FFNLayer (:151-155), decoder_norm + mask_embed_head (models/agile3d.py:342-347) and the query assembly of
 * a click round (models/agile3d.py:202-264).  All matrices [n_scenes, nq, 128] fp32 row-major; `blob` is the
 * per-layer weight blob of ag3d_query_blob_floats() floats (layout: csrc/query_ops.cu, built by
 * agile3d_b200/model.py::_layer_blob from the state_dict); nq <= 256.
 *   ag3d_query_init:     row r of (queries, qpos): src_row[r] >= 0 -> feats[feat_row[r]] (feat_row == NULL: src_row; the
 *                        features may live in another row order than xyz) and fourier(xyz[src_row[r]];
 *                        range of scene scene_of_row[r]) + time_table[time_idx[r]];  src_row[r] = -(k+1) -> learned
 *                        background query k (bg_feat[k], bg_pos[k]).
 *   ag3d_query_fold_c2s: qfold[(h,q),:] = Wk_h^T ((Wq_h (Q+qpos) + bq_h) / 4)            [n_scenes, 8 nq, 128]
 *   ag3d_query_update_a: q1 = LN(Q + out_proj(per-head Wv ctx + bv)) (tail of c2s);  qh/kh/vh = c2c projections
 *   ag3d_query_update_b: q2 = LN(q1 + c2c attention), q3 = LN(q2 + FFN(q2)) -> queries of the next layer;
 *                        A, c, U = folds for ag3d_s2c_mask_fwd;  E = mask_embed_head(decoder_norm(q3))               */
int64_t ag3d_query_blob_floats(void);
int ag3d_query_init(const float* feats, const float* xyz, const float* range, const int32_t* src_row,
                    const int32_t* feat_row, const int32_t* time_idx, const int32_t* scene_of_row, int32_t n_rows, const float* gauss_B,
                    const float* time_table, const float* bg_feat, const float* bg_pos, float* queries, float* qpos,
                    int32_t feats_split /* feats are "split" bf16 hi/lo rows */, ag3d_stream_t stream);
int ag3d_query_fold_c2s(const float* queries, const float* qpos, const float* blob, int32_t n_scenes, int32_t nq,
                        float* qfold, ag3d_stream_t stream);
int ag3d_query_update_a(const float* ctx, const float* queries, const float* qpos, const float* blob, int32_t n_scenes,
                        int32_t nq, float ln_eps, float* q1, float* qh, float* kh, float* vh, ag3d_stream_t stream);
int ag3d_query_update_b(const float* q1, const float* qh, const float* kh, const float* vh, const float* qpos,
                        const float* blob, int32_t n_scenes, int32_t nq, float ln_eps, float* q3, float* A, float* c,
                        float* U, float* E, ag3d_stream_t stream);

/* ---- click -> scene cross-attention (c2s) -------------------------------------------------------------
 * Replaces nn.MultiheadAttention inside CrossAttentionLayer.forward_post as called at
 * models/agile3d.py:283-290 (models/modules/attention_block.py:86-98), with the key/value projections
 * folded into the queries: qfold[(h,q), :] = Wk_h^T q_h / sqrt(d_h), rows h-major.
 *   ctx[(h,q), :] = sum_v softmax_v( qfold[(h,q)] . (x_v + pos_v)  [masked] ) * x_v
 * Mask (models/agile3d.py:365-380): query q of object q_obj[q] is blocked at voxel v iff label[v] != q_obj[q],
 * unless obj_count[q_obj[q]] == 0 (no voxel carries that label -> the row is un-masked).  label == NULL
 * means no mask (first decoder layer).  algo: AG3D_ALGO_SIMT = fp32 FFMA kernel, AG3D_ALGO_TC (= AUTO) = tcgen05
 * flash-decoding kernel (bf16x3, online softmax, context accumulators resident in TMEM).
 * lse (nullable, [heads*nq]) receives the log-sum-exp of every row's scores (+inf for a row with no admissible
 * voxel) — what ag3d_c2s_attn_bwd needs to rebuild the probabilities.                                          */
size_t ag3d_c2s_workspace_bytes(int32_t nq, int32_t heads);
int ag3d_c2s_attn_fwd(const float* x, const float* pos, int64_t nv, const float* qfold, int32_t nq,
                      int32_t heads, const uint8_t* label, const int32_t* q_obj, const int32_t* obj_count,
                      float* ctx, float* lse, int32_t algo, void* ws, size_t ws_bytes, ag3d_stream_t stream);
/* Same operator on "split" rows (x_split / pos_split: [nv, 128] rows of 4 x (64 B bf16 hi | 64 B bf16 lo), what
 * ag3d_pack_split and the tensor-core backbone write): the voxel tiles go from HBM to the tensor core through the TMA
 * engine (2-D box loads into SWIZZLE_128B tiles) with no thread touching them; scores are Qf.x^T + Qf.pos^T.   */
int ag3d_c2s_attn_fwd_split(const float* x_split, const float* pos_split, int64_t nv, const float* qfold, int32_t nq,
                            int32_t heads, const uint8_t* label, const int32_t* q_obj, const int32_t* obj_count,
                            float* ctx, float* lse, void* ws, size_t ws_bytes, ag3d_stream_t stream);

/* ---- scene -> click cross-attention + LayerNorm + mask head (s2c) --------------------------------------
 * Replaces CrossAttentionLayer.forward_post as called at models/agile3d.py:305-312 plus
 * Agile3d.mask_module (models/agile3d.py:342-384), one pass over the voxels:
 *   S[v,(h,q)] = (x_v + pos_v) . A[(h,q)] + c[(h,q)];  a = softmax over q inside each head
 *   y_v = LayerNorm(x_v + sum_{h,q} a[v,(h,q)] U[(h,q)] + bo)              -> x_out
 *   logits[v, o] = max_{q : q_obj[q] == o} y_v . E[q]  (o = 0 background)    -> logits [nv, n_obj]
 *   label[v] = argmax_o logits[v, o] (first maximum);  obj_count[o] += #voxels labelled o (caller zeroes).
 * x_out may alias x.  nq <= 256: up to 32 queries all score columns of a voxel tile live in TMEM at once
 * (decoder_tc.cu); beyond that the queries are walked in groups of 16 with two-pass softmax statistics
 * (decoder_mq.cu, tensor-core path only).  algo: AG3D_ALGO_SIMT = fp32 FFMA kernel; AG3D_ALGO_TC = three
 * chained tcgen05 GEMMs (bf16x3, fp32 accumulate in TMEM), needs the workspace; AUTO = TC when ws is given.      */
size_t ag3d_s2c_workspace_bytes(int32_t nq);
int ag3d_s2c_mask_fwd(const float* x, const float* pos, int64_t nv, const float* A, const float* c,
                      const float* U, const float* bo, const float* ln_w, const float* ln_b, float ln_eps,
                      const float* E, const int32_t* q_obj, int32_t nq, int32_t heads, int32_t n_obj,
                      float* x_out, float* logits, uint8_t* label, int32_t* obj_count, int32_t algo, void* ws,
                      size_t ws_bytes, ag3d_stream_t stream);
/* Same operator on "split" rows (see ag3d_c2s_attn_fwd_split) for at most 24 click queries: x_split / pos_split tiles
 * arrive through the TMA engine, scores are x.A^T + pos.A^T, the updated features leave as split rows.  x_out_split
 * may alias x_split (in place) and may be NULL (last decoder layer: the features are not read again).        */
int ag3d_s2c_mask_fwd_split(const float* x_split, const float* pos_split, int64_t nv, const float* A, const float* c,
                            const float* U, const float* bo, const float* ln_w, const float* ln_b, float ln_eps,
                            const float* E, const int32_t* q_obj, int32_t nq, int32_t heads, int32_t n_obj,
                            float* x_out_split, float* logits, uint8_t* label, int32_t* obj_count, void* ws,
                            size_t ws_bytes, ag3d_stream_t stream);

/* ======================================================================================================
 * Training step (e): what autograd + MinkowskiEngine's backward kernels + ATen do in
 * the reference's engine.py:119-152 (`losses.backward()`, clip_grad_norm_, optimizer.step()).
 * ====================================================================================================== */

/* ---- BatchNorm with batch statistics (MinkowskiBatchNorm in train mode, models/modules/common.py:20-22) ----
 * ag3d_bn_stats: per-channel mean / inverse std over the n rows (biased variance), and the running-stat update
 *   running = (1 - momentum) * running + momentum * batch  (unbiased variance), running_* nullable.
 * ag3d_bn_apply: y = act((z - mean) * invstd * gamma + beta (+ residual)).
 * ag3d_bn_bwd:   g = dy * (y > 0 if AG3D_RELU);  dbeta = sum g;  dgamma = sum g * xhat;
 *                dz = gamma * invstd * (g - dbeta / n - xhat * dgamma / n);  g_out (nullable) receives g — the
 *                gradient of the residual operand.  dz may alias dy.
 * ag3d_col_sum:  sum over rows (bias gradient of lin_squeeze_head).                                          */
size_t ag3d_colreduce_workspace_bytes(int32_t C);
int ag3d_bn_stats(const float* z, int32_t z_ld, int32_t C, int64_t n, float eps, float momentum, float* running_mean,
                  float* running_var, float* mean, float* invstd, void* ws, size_t ws_bytes, ag3d_stream_t stream);
int ag3d_bn_apply(const float* z, int32_t z_ld, const float* mean, const float* invstd, const float* gamma,
                  const float* beta, const float* residual, int32_t res_ld, int32_t C, int64_t n, int32_t flags,
                  float* y, int32_t y_ld, ag3d_stream_t stream);
int ag3d_bn_bwd(const float* z, int32_t z_ld, const float* y, int32_t y_ld, const float* dy, int32_t dy_ld,
                const float* mean, const float* invstd, const float* gamma, int32_t C, int64_t n, int32_t flags,
                float* dz, int32_t dz_ld, float* g_out, int32_t g_ld, float* dgamma, float* dbeta, void* ws,
                size_t ws_bytes, ag3d_stream_t stream);
int ag3d_col_sum(const float* z, int32_t z_ld, int32_t C, int64_t n, float* sum, void* ws, size_t ws_bytes,
                 ag3d_stream_t stream);

/* ---- sparse convolution backward (MinkowskiEngine ConvolutionBackwardKernelGPU) ----------------------------
 * Data gradient: din[i] = sum_k dout[nbr_t[k][i]] @ W[k]^T is itself a sparse convolution over the transposed
 * map, so ag3d_spconv_bwd_data IS ag3d_spconv_fwd called with (dout, nbr_t, weight_t[k] = W[k_t(k)]^T); for a
 * centred odd kernel nbr_t[k] = nbr[K-1-k] (the table is its own transpose with mirrored offsets), for the
 * kernel-2/stride-2 pair the transposed table is the other one of (ag3d_kernel_map, ag3d_kernel_map_transposed).
 * `residual` adds the gradient arriving over the skip branch in the same epilogue.
 * Weight gradient: dW[k] (+)= sum_o in[nbr[k][o]]^T dout[o]   ([K, cin, cout], fp32 FFMA, deterministic
 * split-row partial sums in the workspace).  nbr == NULL: K = 1, identity map — the plain "X^T dY" contraction,
 * also used for the decoder's query-side gradients.  cin, cout multiples of 4.                               */
int ag3d_spconv_bwd_data(const float* dout, int32_t dout_ld, int32_t cout, const int32_t* nbr_t, int32_t K,
                         int64_t n_in, const float* weight_t, const void* weight_t_tc, int32_t cin,
                         const float* residual, int32_t res_ld, float* din, int32_t din_ld, int32_t algo, void* ws,
                         size_t ws_bytes, ag3d_stream_t stream);
size_t ag3d_spconv_bwd_weight_workspace_bytes(int64_t n_out, int32_t K, int32_t cin, int32_t cout);
int ag3d_spconv_bwd_weight(const float* in, int32_t in_ld, int32_t cin, const int32_t* nbr, int32_t K, int64_t n_out,
                           const float* dout, int32_t dout_ld, int32_t cout, float* dweight, int32_t accumulate,
                           void* ws, size_t ws_bytes, ag3d_stream_t stream);
/* The same gradient on the tensor cores (tcgen05, "bf16x4": all four hi/lo products).  Both operands are "split" rows
 * (AG3D_IN_SPLIT format: 64 B bf16 hi | 64 B bf16 lo per 32-channel slab; ag3d_pack_split converts fp32 rows); the
 * neighbour rows of `in_split` and the rows of `dout_split` are gathered by the TMA engine, so the true row counts of
 * both buffers are part of the call.  cin, cout multiples of 32.                                                    */
int ag3d_pack_split(const float* in, int32_t in_ld, int32_t C, int64_t n, float* out, int32_t out_ld,
                    ag3d_stream_t stream);
int32_t ag3d_spconv_bwd_weight_tc_supported(int32_t K, int32_t cin, int32_t cout);
size_t ag3d_spconv_bwd_weight_tc_workspace_bytes(int64_t n_out, int32_t K, int32_t cin, int32_t cout);
int ag3d_spconv_bwd_weight_tc(const float* in_split, int64_t n_in, int32_t in_ld, int32_t cin, const int32_t* nbr,
                              int32_t K, int64_t n_out, const float* dout_split, int32_t dout_ld, int32_t cout,
                              float* dweight, int32_t accumulate, void* ws, size_t ws_bytes, ag3d_stream_t stream);
/* stem (3 -> 32, probes the hash table like ag3d_stem_conv_fwd): dW[k][ci][co] (+)= feats[src_k(v)][ci] dz[v][co] */
size_t ag3d_stem_bwd_weight_workspace_bytes(int32_t ksize);
int ag3d_stem_bwd_weight(const int32_t* coords, const float* feats, int64_t n, const void* table, int64_t cap,
                         int32_t ksize, const float* dz, int32_t dz_ld, float* dweight, int32_t accumulate, void* ws,
                         size_t ws_bytes, ag3d_stream_t stream);
int ag3d_stem_bwd_weight_bricks(const int32_t* coords, const float* feats, int64_t n, const void* table2, int64_t cap2,
                                const int32_t* brick_rows, int32_t ksize, const float* dz, int32_t dz_ld, float* dweight,
                                int32_t accumulate, void* ws, size_t ws_bytes, ag3d_stream_t stream);

/* ---- decoder backward -------------------------------------------------------------------------------------
 * Query-side matrices are row-padded with zeros to hqp = ag3d_decoder_bwd_rows(nq, heads) (head, query) rows
 * and 32 queries; "t" suffix = transposed copy ([128, hqp] / [128, 32]).
 * c2s: rowobj[r] = object a row is restricted to, -1 = unrestricted, -2 = padding / dead row;
 *      dr[r] = dctx[r] . ctx[r].  Writes dx [nv,128] and ds_out [nv,hqp] (dS; dqfold = dS^T (x + pos)).
 * s2c: dxo (nullable) = gradient of x_out, dlogits (nullable) = gradient of the logits [nv, n_obj] (routed to the
 *      first maximal query of every object, as torch.max does).  Writes dx and the per-voxel factors a_out, ds_out
 *      [nv,hqp], dy_out [nv,128], g_out [nv,32]; colsums = [dbo 128 | dln_w 128 | dln_b 128 | dc hqp].
 *      dA = dS^T (x + pos), dU = a^T dy, dE = g^T x_out  via ag3d_spconv_bwd_weight.                          */
int32_t ag3d_decoder_bwd_rows(int32_t nq, int32_t heads);
int ag3d_c2s_attn_bwd(const float* x, const float* pos, int64_t nv, const float* qf, const float* qft,
                      const float* dctx, const float* dctxt, const float* lse, const float* dr, const int32_t* rowobj,
                      int32_t hqp, const uint8_t* label, float* dx, float* ds_out, ag3d_stream_t stream);
/* Point-wise middle of the same backward when its four GEMMs (S = (x+pos) qf^T, dP = x dctx^T, dx = P dctx + dS qf)
 * run as 1x1 tensor-core convolutions (ag3d_spconv_fwd_rows): P = masked exp(S - lse) and dS = P * (dP - dr), in
 * place (s_p: S -> P, dp_ds: dP -> dS; both f32 [nv, hq], hq % 4 == 0).                                            */
int ag3d_c2s_bwd_pointwise(float* s_p, float* dp_ds, const float* lse, const float* dr, const int32_t* rowobj,
                           const uint8_t* label, int64_t nv, int32_t hq, ag3d_stream_t stream);
/* Row-wise pieces of the scene -> click backward (ag3d_s2c_mask_bwd) for the variant whose GEMMs run as 1x1
 * tensor-core convolutions (agile3d_b200/ops.py::s2c_mask_bwd_tc): per-head softmax over the queries (in place on
 * S [nv, hqp]), dS = a (da - <a, da>) per head (in place on da), LayerNorm statistics (n = normalised rows, rstd),
 * LayerNorm backward with the column sums [dbo | dln_w | dln_b], and the first-maximum routing of dlogits.          */
int ag3d_s2c_softmax_heads(float* s_a, int64_t nv, int32_t heads, int32_t nq, int32_t hqp, ag3d_stream_t stream);
int ag3d_s2c_ds(const float* a, float* da_ds, int64_t nv, int32_t heads, int32_t nq, int32_t hqp, ag3d_stream_t stream);
int ag3d_ln_fwd_stats(const float* y, int64_t nv, float eps, float* n_out, float* rstd_out, ag3d_stream_t stream);
size_t ag3d_ln_bwd_workspace_bytes(void);
int ag3d_ln_bwd(const float* t, const float* n, const float* rstd, const float* ln_w, int64_t nv, float* dy,
                float* colsums, void* ws, size_t ws_bytes, ag3d_stream_t stream);
int ag3d_s2c_route(const float* G, const float* dlogits, const int32_t* q_obj, int32_t nq, int32_t n_obj, int64_t nv,
                   float* g_out, ag3d_stream_t stream);
/* ag3d_s2c_route for up to 256 queries: G and g_out are [nv, ld] (ld >= nq, padding columns of g_out are zeroed),
 * arg_ws is an int32 [nv, n_obj] scratch.                                                                    */
int ag3d_s2c_route_ld(const float* G, int32_t ld, const float* dlogits, const int32_t* q_obj, int32_t nq, int32_t n_obj,
                      int64_t nv, int32_t* arg_ws, float* g_out, ag3d_stream_t stream);
size_t ag3d_s2c_bwd_workspace_bytes(int32_t hqp);
int ag3d_s2c_mask_bwd(const float* x, const float* pos, int64_t nv, const float* A, const float* At, const float* c,
                      const float* U, const float* Ut, const float* bo, const float* ln_w, const float* ln_b,
                      float ln_eps, const float* E, const float* Et, const int32_t* q_obj, int32_t nq, int32_t heads,
                      int32_t n_obj, int32_t hqp, const float* dxo, const float* dlogits, float* dx, float* a_out,
                      float* ds_out, float* dy_out, float* g_out, float* colsums, void* ws, size_t ws_bytes,
                      ag3d_stream_t stream);

/* ---- loss (models/criterion.py:84-132) and click loss weights (utils/seg.py:62-89) --------------------------
 * Per scene: sums[0] = sum_v w_v CE(logits_v, t_v), sums[2] = sum_v w_v dice_v (sums is float[4]; [1],[3] = 0);
 * the caller divides by n.  ag3d_loss_bwd: dlogits = w_v / n * (g[0] dCE + g[1] ddice), g on the device.        */
size_t ag3d_loss_workspace_bytes(void);
int ag3d_loss_fwd(const float* logits, int32_t C, int64_t n, const int32_t* target, const float* w, float eps,
                  float* sums, void* ws, size_t ws_bytes, ag3d_stream_t stream);
int ag3d_loss_bwd(const float* logits, int32_t C, int64_t n, const int32_t* target, const float* w, float eps,
                  const float* g, float* dlogits, ag3d_stream_t stream);
/* target[v] outside [0, C) (the reference asserts this range; e.g. the dataset ignore id -1) yields NaN sums / NaN
 * gradient rows instead of an out-of-bounds read: the failure is loud without a host synchronisation.           */
int ag3d_click_loss_weights(const float* xyz, int64_t n, const float* clicks, int32_t n_clicks, float alpha, float beta,
                            float tita, float* w, ag3d_stream_t stream);

/* ---- clip_grad_norm_ + AdamW over flat fp32 buffers (engine.py:146-150) ---------------------------------------
 * ag3d_grad_norm: norm_out[0] = ||g||_2.  ag3d_adamw_step: g is scaled by min(1, max_norm / (norm + 1e-6)) when
 * grad_norm is given (device pointer), then the decoupled-weight-decay Adam update; `step` counts from 1.        */
size_t ag3d_grad_norm_workspace_bytes(void);
int ag3d_grad_norm(const float* g, int64_t n, float* norm_out, void* ws, size_t ws_bytes, ag3d_stream_t stream);
int ag3d_adamw_step(float* p, const float* g, float* m, float* v, int64_t n, float lr, float beta1, float beta2,
                    float eps, float weight_decay, int32_t step, const float* grad_norm, float max_norm,
                    ag3d_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* AGILE3D_B200_H_ */
