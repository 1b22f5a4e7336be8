/*
 * p3d_b200.h -- C ABI of libp3d_b200.so: hand-written sm_100a CUDA kernels for
 * the per-layer sampling / grouping / neighbour-aggregation path of
 * torch-points3d (leo-stan/torch-points3d).
 *
 * Contract
 *   - plain pointers and sizes; every pointer is a DEVICE pointer unless named h_*;
 *   - the caller owns all memory (inputs, outputs, workspaces); the library never
 *     allocates, frees, synchronises the device or changes the current device;
 *   - `stream` is a cudaStream_t passed as void* (0 = legacy default stream);
 *   - all tensors are contiguous, fp32 points/features, int64 indices unless stated;
 *   - return value 0 = launched; <0 = error, text in p3d_last_error() (thread local);
 *   - there is NO CPU fallback: without a CUDA device every entry point fails.
 *
 * Each entry point cites the reference interface it replaces.  The reference is
 * pure Python; its FFI for this path is the pybind11 surface of two wheels
 * (torch-points-kernels 0.7.0 `torch_points_kernels.points_cuda`, torch-cluster
 * 1.5.9 `torch.ops.torch_cluster`) reached from the call sites cited below, plus
 * the pure-torch KPConv_ops.  INTEGRATION.md shows the binding a maintainer adds.
 *
 * Floating-point rule: squared distances are fp32 ((dx*dx + dy*dy) + dz*dz),
 * never contracted into FMAs -> indices are bit-exact
 * against oracle/p3d_oracle.c.
 */
#ifndef P3D_B200_H
#define P3D_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define P3D_OK 0
#define P3D_ERR_INVALID -1   /* bad argument (shape, null pointer, unsupported size) */
#define P3D_ERR_CUDA -2      /* CUDA runtime reported an error at launch */
#define P3D_ERR_WORKSPACE -3 /* workspace missing or too small */

#define P3D_INFLUENCE_CONSTANT 0
#define P3D_INFLUENCE_LINEAR 1
#define P3D_INFLUENCE_GAUSSIAN 2
#define P3D_AGGREGATION_SUM 0
#define P3D_AGGREGATION_CLOSEST 1
#define P3D_PRECISION_FP32 0   /* exact fp32 SIMT arithmetic (1e-5 relative against the oracle) */
#define P3D_PRECISION_TENSOR 1 /* kernel-point contraction on tcgen05, bf16 hi+lo split operands, fp32 accumulate */

int p3d_version(void);
const char *p3d_last_error(void);
/* number of kernels this library has launched so far in this process (bench.py: gpu_launches) */
uint64_t p3d_launch_count(void);
/* compiled-for architecture string, e.g. "sm_100a" */
const char *p3d_arch(void);

/* ---------------------------------------------------------------- FPS ------
 * Replaces tp.furthest_point_sample(xyz[B,N,3], npoint) -> idx[B,npoint]
 *   reference call sites: torch_points3d/core/spatial_ops/sampling.py:100,
 *                         torch_points3d/modules/VoteNet/proposal_module.py:57
 * Start index 0, lowest-index tie-break.  idx_is_i64: 0 -> int32 out (what tpk's CUDA
 * path returns), 1 -> int64 out (what core/base_conv/dense.py:75 casts to).
 * N <= 131072 runs register-resident in one CTA / one thread-block cluster per cloud;
 * larger N needs `ws` = B*N floats (min-distance scratch).
 * new_xyz (nullable) [B,S,3]: the sampled coordinates, i.e. the `pos.gather(1, idx)` of
 * core/base_conv/dense.py:75-76 for free (every round already knows its winner's coordinates). */
size_t p3d_fps_dense_ws_bytes(int B, int N);
int p3d_fps_dense_f32(const float *xyz, int B, int N, int S, void *idx, int idx_is_i64, float *new_xyz, void *ws,
                      size_t ws_bytes, void *stream);

/* Replaces torch_geometric.nn.fps(x, batch, ratio, random_start=False) == torch_cluster.fps
 *   reference call sites: core/spatial_ops/sampling.py:63,
 *                         datasets/registration/utils.py:337, modules/VoteNet/dense_samplers.py:70,93
 * CSR form: cloud c owns rows ptr[c]..ptr[c+1] and emits out_ptr[c+1]-out_ptr[c] GLOBAL
 * indices at idx[out_ptr[c]..].  start (nullable): per-cloud LOCAL index of the first sample
 * (random_start=True draws it on the host side; NULL = index 0, the pinned behaviour of
 * test/test_fps.py:35-42).  max_n = host-known upper bound on the largest cloud (selects the
 * register tiling); ws as above with total rows. */
size_t p3d_fps_csr_ws_bytes(int64_t total_rows, int64_t max_n);
int p3d_fps_csr_f32(const float *xyz, const int64_t *ptr, int num_clouds, const int64_t *out_ptr, int64_t *idx,
                    const int64_t *start, int64_t max_n, void *ws, size_t ws_bytes, void *stream);

/* --------------------------------------------------------- ball query ------
 * Replaces tp.ball_query(radius, nsample, x[B,N,3], y[B,S,3], mode="dense") -> (idx, dist2)
 *   reference call site: core/spatial_ops/neighbour_finder.py:164
 * First `nsample` candidates in index order with d2 < r2 (strict != 0) or d2 <= r2;
 * padding = first hit, empty ball = 0; dist2 (nullable) padded with -1. */
int p3d_ball_query_dense_f32(const float *x, const float *y, int B, int N, int S, float radius, int nsample,
                             int strict, int64_t *idx, float *dist2, void *stream);

/* All scales of a multi-scale-grouping layer in one sweep: DenseRadiusNeighbourFinder is called once per scale on
 * the same (x, y) pair (core/base_conv/dense.py:78-82 over neighbour_finder.py:156-178; PointNet++ MSG uses 2-3
 * radii per level).  Every squared distance is computed once and tested against each radius; idx of scale s is
 * bit-identical to p3d_ball_query_dense_f32(radius[s], nsample[s]).  h_radii, h_nsamples, h_idx are HOST arrays of
 * num_scales (<= 4) entries; h_idx[s] is the device pointer of idx_s [B,S,nsample[s]]. */
int p3d_ball_query_dense_multi_f32(const float *x, const float *y, int B, int N, int S, int num_scales,
                                   const float *h_radii, const int *h_nsamples, int strict, int64_t *const *h_idx,
                                   void *stream);

/* Replaces tp.ball_query(..., mode="partial_dense", batch_x, batch_y)
 *   reference call site: core/spatial_ops/neighbour_finder.py:35-37
 * batch_x[Nx], batch_y[Ny]: sorted cloud ids (int64).  GLOBAL indices into x, unfilled
 * slots = shadow (the reference uses x.shape[0]), dist2 (nullable) = -1.  The cloud
 * ranges are found on the device (no blocking D2H copy of batch_x[-1]). */
int p3d_ball_query_batch_f32(const float *x, const float *y, const int64_t *batch_x, const int64_t *batch_y,
                             int64_t Nx, int64_t Ny, float radius, int nsample, int strict, int64_t shadow,
                             int64_t *idx, float *dist2, void *stream);

/* ---------------------------------------------------------------- knn ------
 * Replaces torch_geometric.nn.knn(x, y, k, batch_x, batch_y) == torch_cluster.knn
 *   reference call sites: core/spatial_ops/neighbour_finder.py:47, interpolate.py:25,70
 * col[Ny,k]: GLOBAL support indices, ascending d2, ties -> lower index, -1 where the cloud
 * has fewer than k supports; d2 (nullable) likewise (-1).  batch_* nullable = one cloud.
 * k <= 128. */
int p3d_knn_batch_f32(const float *x, const float *y, const int64_t *batch_x, const int64_t *batch_y, int64_t Nx,
                      int64_t Ny, int k, int64_t *col, float *d2, void *stream);

/* Dense-layout knn: tp.knn(pos_support[B,N,3], pos[B,M,3], k) -> (idx[B,M,k], dist2[B,M,k])
 *   reference call sites: core/data_transform/transforms.py:16 (dense_knn), datasets/registration/*
 * indices are local to the cloud; k <= N required by the caller. */
int p3d_knn_dense_f32(const float *x, const float *y, int B, int N, int M, int k, int64_t *idx, float *d2,
                      void *stream);

/* Cell-list accelerated forms of the two batched searches (csrc/cellgrid.cu).  Same arguments and BIT-IDENTICAL
 * results as p3d_ball_query_batch_f32 / p3d_knn_batch_f32 -- first `nsample` hits in candidate index order, resp. the k
 * nearest ordered by (d2, index) -- but a query only meets the support points of the 27 lattice cells around it
 * (knn: growing cubes of cells, with a plain sweep when the neighbourhood is sparse).  The support points are hashed
 * by (cloud, cell) on the device in `ws` (p3d_cellgrid_ws_bytes(Nx) bytes, 256-byte aligned); nothing is read back.
 * Worth it when clouds hold thousands of points (KPConv levels: neighbour_finder.py:34-37 at ~6k points per cloud;
 * RandLA / FPModule knn: neighbour_finder.py:46-47, interpolate.py:25); nsample <= 64, k <= 32. */
size_t p3d_cellgrid_ws_bytes(int64_t Nx);
int p3d_ball_query_batch_grid_f32(const float *x, const float *y, const int64_t *batch_x, const int64_t *batch_y,
                                  int64_t Nx, int64_t Ny, float radius, int nsample, int strict, int64_t shadow,
                                  int64_t *idx, float *dist2, void *ws, size_t ws_bytes, void *stream);
int p3d_knn_batch_grid_f32(const float *x, const float *y, const int64_t *batch_x, const int64_t *batch_y, int64_t Nx,
                           int64_t Ny, int k, int64_t *col, float *d2, void *ws, size_t ws_bytes, void *stream);


/* Replaces tp.three_nn(unknown[B,n,3], known[B,m,3]) -> (dist[B,n,3] = sqrt(d2), idx[B,n,3])
 *   reference call site: core/base_conv/dense.py:136
 * weight (nullable) [B,n,3]: the inverse-distance weights of core/base_conv/dense.py:137-139,
 * (1/(d+1e-8)) / sum_t (1/(d_t+1e-8)), fused into the same kernel. */
int p3d_three_nn_f32(const float *unknown, const float *known, int B, int n, int m, float *dist, int64_t *idx,
                     float *weight, void *stream);

/* ----------------------------------------------------------- grouping ------
 * Replaces tp.grouping_operation(features[B,C,N], idx[B,S,K]) -> [B,C,S,K]
 *   reference call sites: modules/pointnet2/dense.py:38,45; modules/RSConv/dense.py:147,152,...
 * out is addressed as out[b*out_batch_stride + (c_off + c)*S*K + s*K + k] so that the
 * caller can write straight into a concatenated [B,3+C,S,K] tensor
 * (out_batch_stride = C*S*K, c_off = 0 for the plain op). */
int p3d_group_fwd_f32(const float *feat, const int64_t *idx, int B, int C, int N, int S, int K, float *out,
                      int64_t out_batch_stride, int c_off, void *stream);
/* backward (autograd of the gather): dfeat[b,c,i] = sum_{(s,k): idx=i} dout[...]; dfeat is
 * overwritten.  dout is addressed like `out` above.  With a workspace of
 * p3d_group_bwd_ws_bytes() the sum is atomic-free and deterministic (the index tensor is
 * inverted by a stable counting sort, slots are summed in increasing (s,k) order); source tables
 * of more than 4096 points per cloud report 0 workspace bytes and use fp32 reductions at L2. */
size_t p3d_group_bwd_ws_bytes(int B, int C, int N, int S, int K);
/* Optional: build the inverse index into `ws` ahead of time (it depends on idx only, so the forward pass
 * can run it on a side stream); pass the same ws with prepared = 1 to p3d_group_bwd_f32 afterwards. */
int p3d_group_bwd_prepare(const int64_t *idx, int B, int C, int N, int S, int K, void *ws, size_t ws_bytes,
                          void *stream);
int p3d_group_bwd_f32(const float *dout, const int64_t *idx, int B, int C, int N, int S, int K, float *dfeat,
                      int64_t dout_batch_stride, int c_off, void *ws, size_t ws_bytes, int prepared, void *stream);
/* Fused PointNetMSGDown._prepare_features xyz part (modules/pointnet2/dense.py:37-42):
 * out[b, c, s, k] = (pos[b, idx[b,s,k], c] - new_pos[b, s, c]) * (1/norm)   for c in 0..2,
 * norm = 1 unless normalize_xyz (ATen evaluates `tensor /= python_float` on CUDA as a multiply
 * by the fp32 reciprocal; the kernel does the same).  pos is the untransposed [B,N,3] tensor. */
int p3d_group_xyz_f32(const float *pos, const float *new_pos, const int64_t *idx, int B, int N, int S, int K,
                      float norm, float *out, int64_t out_batch_stride, void *stream);

/* Both halves of _prepare_features in ONE launch: out [B,3+C,S,K] = cat([(pos[idx] - new_pos) / norm, feat[idx]], 1)
 * (modules/pointnet2/dense.py:36-54).  The relative-coordinate channels are one more channel block of the staged
 * gather kernel (positions transposed into shared memory once per CTA).  Tables too large to stage (N > 8192) run
 * the two kernels above. */
int p3d_group_xyz_feat_f32(const float *pos, const float *new_pos, const float *feat, const int64_t *idx, int B, int C,
                           int N, int S, int K, float norm, float *out, void *stream);

/* ---------------------------------------------------- three_interpolate ----
 * Replaces tp.three_interpolate(features[B,C,m], idx[B,n,3], weight[B,n,3]) -> [B,C,n]
 *   reference call site: core/base_conv/dense.py:140 */
int p3d_three_interp_fwd_f32(const float *feat, const int64_t *idx, const float *weight, int B, int C, int m,
                             int n, float *out, void *stream);
/* backward w.r.t. features; dfeat is overwritten; workspace and determinism as for
 * p3d_group_bwd_f32 (m <= 4096 known points per cloud, else fp32 reductions at L2) */
size_t p3d_three_interp_bwd_ws_bytes(int B, int C, int m, int n);
int p3d_three_interp_bwd_prepare(const int64_t *idx, int B, int C, int m, int n, void *ws, size_t ws_bytes,
                                 void *stream);
int p3d_three_interp_bwd_f32(const float *dout, const int64_t *idx, const float *weight, int B, int C, int m,
                             int n, float *dfeat, void *ws, size_t ws_bytes, int prepared, void *stream);

/* ------------------------------------------------------------- KPConv ------
 * Replaces KPConv_ops(query, support, neighbors, features, K_points, K_values, KP_extent,
 *                     KP_influence, aggregation_mode) -> [n, Cout]
 *   reference: torch_points3d/modules/KPConv/convolution_ops.py:19-107
 *   called from modules/KPConv/kernels.py:83-93
 * neighbors[n,M] int64, entries >= N0 or < 0 are shadow neighbours (point at 1e6, zero
 * feature).  K <= 16 kernel points.  One fused kernel: no [n,M,K] / [n,M,Cin] / [n,K,Cin]
 * intermediate is written to HBM.
 * precision = P3D_PRECISION_TENSOR runs the (kernel point x Cin) -> Cout contraction on tcgen05
 * tensor cores (TMA-fed pre-split weights, fp32 accumulation in TMEM; operands are bf16 hi+lo pairs,
 * three partial products (hi.hi, lo.hi, hi.lo) per product: relative error ~1e-5 of max|out|, stated and tested at 2e-5) when
 * Cin % 32 == 0, Cout % 32 == 0, Cout <= 512, M <= 40 and a workspace of
 * p3d_kpconv_fwd_ws_bytes() is given; other shapes and P3D_PRECISION_FP32 use exact fp32 SIMT. */
size_t p3d_kpconv_fwd_ws_bytes(int64_t n, int K, int Cin, int Cout);
int p3d_kpconv_fwd_f32(const float *query, const float *support, const int64_t *neighbors, const float *feat,
                       const float *k_points, const float *k_values, int64_t n, int64_t N0, int M, int K, int Cin,
                       int Cout, float extent, int influence, int aggregation, int precision, float *out,
                       void *ws, size_t ws_bytes, void *stream);
/* autograd backward: dfeat[N0,Cin] (zero-initialised by the caller, nullable) and
 * dweights[K,Cin,Cout] (zero-initialised by the caller, nullable).  ws: p3d_kpconv_bwd_ws_bytes. */
size_t p3d_kpconv_bwd_ws_bytes(int64_t n, int M, int K, int Cin, int Cout);
int p3d_kpconv_bwd_f32(const float *query, const float *support, const int64_t *neighbors, const float *feat,
                       const float *k_points, const float *k_values, const float *dout, int64_t n, int64_t N0,
                       int M, int K, int Cin, int Cout, float extent, int influence, int aggregation,
                       float *dfeat, float *dweights, void *ws, size_t ws_bytes, void *stream);

/* The geometric halves of KPConv_ops on their own (convolution_ops.py:49-98 and its adjoint), for shapes where the
 * dense contraction is better run as a plain GEMM by the caller (few query points, wide channels):
 *   wf[n,K,Cin] = sum_m influence(p,m,k) * feat[neighbors[p,m]]            out = wf.view(n, K*Cin) @ K_values.view(K*Cin, Cout)
 *   dfeat[neighbors[p,m]] += sum_k influence(p,m,k) * dwf[p,k,:]           (dfeat zeroed by the caller) */
int p3d_kpconv_wf_f32(const float *query, const float *support, const int64_t *neighbors, const float *feat,
                      const float *k_points, int64_t n, int64_t N0, int M, int K, int Cin, float extent, int influence,
                      int aggregation, float *wf, void *stream);
int p3d_kpconv_scatter_f32(const float *query, const float *support, const int64_t *neighbors, const float *k_points,
                           const float *dwf, int64_t n, int64_t N0, int M, int K, int Cin, float extent, int influence,
                           int aggregation, float *dfeat, void *stream);

/* ------------------------------------------------------ deformable KPConv ------
 * Replaces KPConv_deform_ops(query, support, neighbors, features, K_points, offsets, modulations, K_values,
 *                            KP_extent, KP_influence, "sum") -> (out [n,Cout], sq_distances [n,M,K], K_points + offsets)
 *   reference: torch_points3d/modules/KPConv/convolution_ops.py:110-235, called from
 *   KPConvDeformableLayer.forward (modules/KPConv/kernels.py:228-241).
 * offsets [n,K,3]; modulations (nullable) [n,K].  Neighbours in range (d2 < extent^2) of no deformed kernel point are
 * dropped (the reference's topk re-compaction).  aggregation_mode "closest" cannot run in the reference (it passes a
 * TensorFlow keyword to torch.argmin, :196) and is not offered.
 * wf [n,K,Cin] (required): modulated weighted features, out = wf . K_values; keep it for the backward pass.
 * sq_distances (nullable) [n,M,K]; kpmin (nullable) [n,K] + argmin (nullable) [n,K] int32: min over the neighbours of
 * sq_distances, which is all fitting_loss (modules/KPConv/losses.py:4-16) reads -- saves writing and re-reading the
 * [n,M,K] tensor.  deformed_K_points is K_points + offsets (caller side). */
int p3d_kpconv_deform_fwd_f32(const float *query, const float *support, const int64_t *neighbors, const float *feat,
                              const float *k_points, const float *offsets, const float *modulations,
                              const float *k_values, int64_t n, int64_t N0, int M, int K, int Cin, int Cout,
                              float extent, int influence, float *out, float *wf, float *sq_distances, float *kpmin,
                              int32_t *argmin, void *stream);
/* autograd backward.  Cotangents: dout [n,Cout]; d_sq_distances (nullable) [n,M,K]; d_kpmin (nullable) [n,K] with the
 * forward's argmin.  Gradients (all nullable): dfeat [N0,Cin] and dweights [K,Cin,Cout] are ACCUMULATED into (zero them
 * first); doffsets [n,K,3] and dmodulations [n,K] are overwritten (the cotangent of deformed_K_points adds to doffsets
 * on the caller side). */
size_t p3d_kpconv_deform_bwd_ws_bytes(int64_t n, int K, int Cin);
int p3d_kpconv_deform_bwd_f32(const float *query, const float *support, const int64_t *neighbors, const float *feat,
                              const float *k_points, const float *offsets, const float *modulations,
                              const float *k_values, const float *wf, const float *dout, const float *d_sq_distances,
                              const float *d_kpmin, const int32_t *argmin, int64_t n, int64_t N0, int M, int K, int Cin,
                              int Cout, float extent, int influence, float *dfeat, float *dweights, float *doffsets,
                              float *dmodulations, void *ws, size_t ws_bytes, void *stream);

/* Training pair: the tensor-path forward also stores the fp32 weighted features wf [n,K,Cin] it holds in registers
 * (one extra 4*K*Cin-byte row per point), and the backward reads them for dW = wf^T dout instead of recomputing the
 * influence-weighted aggregation (a third of the backward at the shallow levels).  p3d_kpconv_fwd_wf_f32 fails with
 * P3D_ERR_INVALID for shapes outside the tensor path; wf = NULL in p3d_kpconv_bwd_wf_f32 recomputes. */
int p3d_kpconv_fwd_wf_f32(const float *query, const float *support, const int64_t *neighbors, const float *feat,
                          const float *k_points, const float *k_values, int64_t n, int64_t N0, int M, int K, int Cin,
                          int Cout, float extent, int influence, int aggregation, float *out, float *wf, void *ws,
                          size_t ws_bytes, void *stream);
int p3d_kpconv_bwd_wf_f32(const float *query, const float *support, const int64_t *neighbors, const float *feat,
                          const float *k_points, const float *k_values, const float *dout, const float *wf, int64_t n,
                          int64_t N0, int M, int K, int Cin, int Cout, float extent, int influence, int aggregation,
                          float *dfeat, float *dweights, void *ws, size_t ws_bytes, void *stream);

/* ---- tall GEMM of the KPConv blocks -----------------------------------------------------------------------------
 * C[R,N] (row stride ldc) = A[R,Kd] (row stride lda) . B, B(k,j) = B[k*sbk + j*sbj], fp32 in and out, on the tcgen05
 * tensor cores (bf16 hi+lo operands, the three significant partial products, fp32 accumulation in TMEM: error
 * <= 2e-5 of max|C|).  R = points of a level, Kd / N = channel counts.  It is the product behind the unary layers of
 * the KPConv blocks (modules/KPConv/blocks.py:176-194: y = x . W^T -> sbk = 1, sbj = Cin; dx = dy . W -> sbk = Cin,
 * sbj = 1) and behind the KPConv backward's dwf = dout . W^T.  N must be a multiple of 16 (else P3D_ERR_INVALID);
 * ws: p3d_rowgemm_ws_bytes(Kd, N) bytes, 16-byte aligned. */
size_t p3d_rowgemm_ws_bytes(int Kd, int N);
int p3d_rowgemm_f32(const float *A, int64_t lda, const float *B, int64_t sbk, int64_t sbj, int64_t R, int Kd, int N,
                    float *C, int64_t ldc, void *ws, size_t ws_bytes, void *stream);
/* dW[R,N] += X^T . B over the n rows of X[n,R] and B[n,N] (the weight gradient of such a layer: X = dy, B = x gives
 * dW[Cout,Cin]; the KPConv weight gradient: X = weighted features, B = dout).  Same arithmetic; the reduction is cut
 * over CTAs and the partial outputs are added in a fixed order (deterministic).  dW is accumulated into (zero it
 * first for a plain product).  N must be a multiple of 16; ws: p3d_rowgemm_dw_ws_bytes(n, R, N). */
size_t p3d_rowgemm_dw_ws_bytes(int64_t n, int R, int N);
int p3d_rowgemm_dw_f32(const float *X, const float *B, int64_t n, int R, int N, float *dW, void *ws, size_t ws_bytes,
                       void *stream);

/* Strided-shortcut neighbour max-pool (modules/KPConv/blocks.py:206-210):
 * out[p,c] = max_m x'[neighbors[p,m], c], x' = cat(x, zeros); arg (nullable) = winning row
 * (N0 for the shadow row) for the backward pass. */
int p3d_neighbor_max_f32(const float *x, const int64_t *neighbors, int64_t n, int64_t N0, int M, int C,
                         float *out, int64_t *arg, void *stream);
/* its backward: dx[N0,C] (overwritten) = sum over the query points whose arg[p,c] names the row, of dout[p,c] */
int p3d_neighbor_max_bwd_f32(const float *dout, const int64_t *arg, int64_t n, int64_t N0, int C, float *dx,
                             void *stream);

/* ------------------------------------------------------ GridSampling3D ------
 * Replaces the position side of GridSampling3D._process (core/data_transform/grid_transform.py:110-126):
 *     coords = torch.round(pos / size); cluster = voxel_grid(coords, batch, 1)   [torch_cluster.grid_cluster]
 *     cluster, unique_pos_indices = consecutive_cluster(cluster)                 [torch_geometric 1.7.2]
 *   reference call sites: modules/KPConv/blocks.py:74-77 (strided blocks), core/data_transform/transforms.py:603-651
 *   (MultiScaleTransform), core/spatial_ops/sampling.py:75 (RandomSampler/GridSampler).
 * batch (nullable) [N] int64.  Voxels are numbered in increasing (batch, z, y, x) order of their rounded
 * coordinates (grid_cluster strides: x fastest, batch slowest; sorted unique ids).  Outputs:
 *   cluster[N]      consecutive voxel id of every point;
 *   last_index[N]   first n rows: highest point index inside voxel c (what consecutive_cluster's sequential
 *                   scatter_ leaves in `perm`; group_data mode="last" and the `batch` key select these rows);
 *   order[N], seg_start[N+1]  points sorted by (voxel, point index) and the first slot of every voxel
 *                   (seg_start[n] = N): the segment list the two companion entry points below consume;
 *   mean_pos (nullable) [N,3]   first n rows: scatter_mean of pos, fp32 sums in ascending point index;
 *   out_batch (nullable) [N]    first n rows: batch[last_index];
 *   coords_out (nullable) [N,3] int32, first n rows: the rounded coordinates of last_index (quantize_coords=True);
 *   n_out           device scalar: number of voxels n.
 * One cooperative persistent kernel (needs all of its <= 148 CTAs co-resident). N < 2^31. */
size_t p3d_grid_cluster_ws_bytes(int64_t N);
int p3d_grid_cluster_f32(const float *pos, const int64_t *batch, int64_t N, float size, int64_t *cluster,
                         int64_t *last_index, int32_t *order, int32_t *seg_start, float *mean_pos,
                         int64_t *out_batch, int32_t *coords_out, int64_t *n_out, void *ws, size_t ws_bytes,
                         void *stream);
/* scatter_mean(item, cluster, dim=0) of group_data (grid_transform.py:76): out[c,:] = mean of x[p,:] over the
 * voxel's points, summed in ascending point index.  n_rows = rows to produce; n_dev (nullable) = device
 * scalar clamping it (lets a caller launch before it has read n back). */
int p3d_segment_mean_f32(const float *x, const int32_t *order, const int32_t *seg_start, const int64_t *n_dev,
                         int64_t n_rows, int C, float *out, void *stream);
/* majority label per voxel (grid_transform.py:70-74: one_hot(item - min) -> scatter_add -> argmax + min; ties
 * go to the lowest label).  hist [n, num_labels] int32 must be zeroed by the caller. */
int p3d_segment_vote_i64(const int64_t *labels, const int64_t *cluster, int64_t N, int64_t n, int64_t label_min,
                         int num_labels, int32_t *hist, int64_t *out, void *stream);

/* ------------------------------------------------ MESSAGE_PASSING operators ------
 * Edge lists [2,E] come from radius / knn (core/spatial_ops/neighbour_finder.py:32-33,47): src[e] indexes the
 * support set, dst[e] the query set, sorted by dst (BaseConvolutionDown.forward, core/base_conv/message_passing.py:44-58
 * stacks them as edge_index = [col, row]).
 * p3d_edge_message: input of PointConv's local_nn (torch_geometric 1.7.2 PointConv.message, used by
 *   modules/pointnet2/message_passing.py:9-26): out[e] = [x[src[e]] (C, may be 0), pos_src[src[e]] - pos_dst[dst[e]]].
 * p3d_edge_message_bwd: dx[src[e]] += dout[e, 0:C] (dx zeroed by the caller).
 * p3d_edge_encoding: RandlaKernel.message (modules/RandLANet/modules.py:28-40) relative point position encoding
 *   out[e] = [pos_i, pos_j, pos_i - pos_j, |pos_i - pos_j|] (10 floats), i = dst, j = src.
 * p3d_segment_reduce: MessagePassing.aggregate over the contiguous edges of every target t in [0,T):
 *   mode 0 = max (PointConv; arg (nullable) [T,C] = winning edge, -1 and value 0 for a target without edges),
 *   mode 1 = sum in edge order (RandlaKernel).  p3d_segment_reduce_bwd: dmsg [E,C] from dout [T,C]. */
int p3d_edge_message_f32(const float *x, const float *pos_src, const float *pos_dst, const int64_t *src,
                         const int64_t *dst, int64_t E, int C, float *out, void *stream);
int p3d_edge_message_bwd_f32(const float *dout, const int64_t *src, int64_t E, int C, float *dx, void *stream);
int p3d_edge_encoding_f32(const float *pos_src, const float *pos_dst, const int64_t *src, const int64_t *dst, int64_t E,
                          float *out, void *stream);
int p3d_segment_reduce_f32(const float *msg, const int64_t *dst, int64_t E, int64_t T, int C, int mode, float *out,
                           int64_t *arg, void *stream);
int p3d_segment_reduce_bwd_f32(const float *dout, const int64_t *dst, const int64_t *arg, int64_t E, int64_t T, int C,
                               int mode, float *dmsg, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* P3D_B200_H */
