/*
 * lanczosnet_b200.h -- C ABI of liblanczosnet_b200.so (sm_100a only).
 *
 * Drop-in boundary for the LanczosNet spectral-convolution forward path of
 * lrjconan/LanczosNetwork.  Conventions follow the reference's own native interface
 * (operators/src/cuda/segment_reduction.h:8-12): the CUDA stream comes first, inputs are
 * const device pointers to contiguous row-major buffers, dimensions are plain ints,
 * outputs come last.  Differences, on purpose:
 *   - every entry point returns an int status (0 = ok, <0 = argument error, >0 = cudaError_t)
 *     instead of calling exit(-1) on a launch failure (segment_reduction.cu:28-36);
 *   - the library owns no tensor memory and never synchronises; workspace is caller-provided;
 *   - no global mutable state besides a thread-local error string (re-entrant under
 *     nn.DataParallel's one-thread-per-device execution, runner/qm8_runner.py:291-292).
 *
 * All pointers are DEVICE pointers unless stated otherwise.  There is no CPU fallback.
 */
#ifndef LANCZOSNET_B200_H_
#define LANCZOSNET_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* lnb_stream_t; /* cudaStream_t */

#define LNB_OK 0
#define LNB_ERR_ARG (-1)
#define LNB_ERR_UNSUPPORTED (-2)

/* ABI version (bumped on any signature change) and last error text of the calling thread. */
int lnb_abi_version(void);
const char* lnb_last_error(void);
/* Number of kernels this library has launched from the calling thread (bench.py's
 * gpu_launches evidence). */
int64_t lnb_launch_count(void);

/* ---------------------------------------------------------------------------------------
 * operators/segment_reduction  (replaces operators/src/cuda/segment_reduction.h:8-12 and the
 * four python-visible names of operators/src/segment_reduction{,_cuda}.h:1-6)
 * data [B, dim1, dim2] fp32, segment_ids [B, dim1] int64, output [B, num_segments, dim2].
 * forward:  output[b, ids[b,c], :] += data[b, c, :]       (output pre-zeroed by the caller,
 *                                                          operators/functions/unsorted_segment_sum.py:20-27)
 * backward: grad_data[b, c, :] = grad_output[b, ids[b,c], :]
 * data_shape = host pointer to {B, dim1, dim2} like the reference launcher.
 * ------------------------------------------------------------------------------------- */
int lnb_unsorted_segment_sum_forward(lnb_stream_t stream, const float* data,
                                     const int64_t* segment_ids, const int* data_shape,
                                     int num_segments, float* output);
int lnb_unsorted_segment_sum_backward(lnb_stream_t stream, const float* grad_output,
                                      const int64_t* segment_ids, const int* data_shape,
                                      int num_segments, float* grad_data);

/* ---------------------------------------------------------------------------------------
 * Generic strided batched fp32 GEMM
 *     C[b,z] = act( alpha * (A[b,z] * kscale[b,z]) @ B[b,z] + beta * D[b,z] + bias )
 * (the torch.bmm / nn.Linear call sites of model/lanczos_net.py:112-121,167-181; the addend D is
 * the Chebyshev recurrence 2 L s_{k-1} - s_{k-2} of model/cheby_net.py:91-93).
 * Strides are in elements.  kscale (optional) multiplies column k of A; bias (optional) has
 * N entries; relu != 0 applies max(.,0); alpha == 0 is read as 1 (zero-initialised descriptors
 * keep their old meaning); addend == NULL skips the beta term.  CUDA-core FFMA path for
 * arbitrary shapes/strides.
 * ------------------------------------------------------------------------------------- */
typedef struct lnb_gemm_desc {
  const float* A; int64_t a_sb, a_sz, a_sm, a_sk;
  const float* B; int64_t b_sb, b_sz, b_sk, b_sn;
  float* C;       int64_t c_sb, c_sz, c_sm, c_sn;
  const float* kscale; int64_t s_sb, s_sz, s_sk;
  const float* bias; int64_t bias_sz;   /* bias[z*bias_sz + n] */
  int32_t batch, nz, M, N, K, relu;
  float alpha, beta;
  const float* addend; int64_t d_sb, d_sz, d_sm, d_sn;
} lnb_gemm_desc;
int lnb_batched_gemm(lnb_stream_t stream, const lnb_gemm_desc* desc /* host */);

/* ---------------------------------------------------------------------------------------
 * Dense layer on 5th-gen tensor cores:  C[M,N] = act(A[M,K] @ W[N,K]^T + bias)
 * (nn.Linear of model/lanczos_net.py:181 and the 4096-wide MLP of ada_lanczos_net.py:54-63).
 * fp32 in/out; computed as 3xTF32 split products (hi*hi + hi*lo + lo*hi) with fp32 TMEM
 * accumulation -> fp32-grade accuracy.  W_hi / W_lo are the tf32 split of W produced once by
 * lnb_split_tf32.  Requirements: K % 4 == 0 and 16-byte aligned operands; any M, N.
 * Two TMEM accumulators (A_hi*W_hi and the two correction products) are summed in fp32 by the
 * epilogue so the small terms do not add truncation steps to the large accumulator.
 * ------------------------------------------------------------------------------------- */
int lnb_split_tf32(lnb_stream_t stream, const float* x, int64_t n, float* hi, float* lo);
int lnb_linear_tf32x3(lnb_stream_t stream, const float* A, const float* W_hi, const float* W_lo,
                      const float* bias, int M, int N, int K, int relu, float* C);
/* Split-K variant for few output tiles and a deep K (the 4096-wide learned filter MLP of
 * model/ada_lanczos_net.py:54-63 at M = batch): every 128 x 128 output tile is computed by `splits`
 * CTAs over disjoint K ranges, so 2 x 32 tiles fill 128 SMs instead of 64 and the weight stream uses
 * the whole HBM bandwidth.  workspace: ceil(M/128)*ceil(N/128)*splits*128*128 floats; counters:
 * ceil(M/128)*ceil(N/128) ints, zero on entry (the kernel leaves them zero). */
int lnb_linear_tf32x3_splitk(lnb_stream_t stream, const float* A, const float* W_hi, const float* W_lo,
                             const float* bias, int M, int N, int K, int relu, float* C, int splits,
                             float* workspace, int* counters);

/* Block-diagonal ("grouped") variant: C[:, g*N:(g+1)*N] = act(A[:, g*K:(g+1)*K] @ W_g^T + b_g)
 * with A [M, groups*K], W stacked [groups*N, K], bias [groups*N], C [M, groups*N].  Used to run
 * the per-layer Ritz-filter MLPs of all layers (model/lanczos_net.py:47-58,109-113) in one launch
 * per MLP stage. */
int lnb_linear_tf32x3_grouped(lnb_stream_t stream, const float* A, const float* W_hi,
                              const float* W_lo, const float* bias, int M, int groups, int N, int K,
                              int relu, float* C);

/* ---------------------------------------------------------------------------------------
 * Fused spectral graph-convolution layer (model/lanczos_net.py:157-182):
 *   out[b,n,:] = act( cat_c(M_c X_b)[n,:] W^T + bias ),  channels c = S long scales
 *   (V diag(coeff[:,:,s]) V^T) followed by the E1 edge-type operators L[...,e].
 * One persistent tcgen05 kernel; no intermediate of the reference (N x N filters, [B*N, C*D]
 * messages) exists in HBM.  lnb_graph_prepare runs once per forward (the operators are layer
 * invariant):
 *   ell_val/ell_idx [B,E1,N,N]  t-major ELL rows of every operator channel, ell_max [B,E1];
 *   gext [B,2] = {n_eff, k_eff}: operators / Q are identically zero beyond these extents;
 *   tiles [4B+2]: tiles[0] = T, tiles[1+t] = first graph of packed tile t (next-fit:
 *                sum n_eff <= 128, sum ceil4(k_eff) <= 128, <= 32 graphs), tiles[1+T] = B;
 *                entries [B+2, 4B+2) are scratch of the assignment kernel;
 *   rowmap [B*K], nrows [1] (both optional, NULL to skip): the compact Ritz row list of
 *                lnb_ritz_rowmap, produced by the same pass;
 *   flags bit 0: store 1.0 for every non-zero (the `L[L != 0] = 1.0` of model/gcnfp.py:83).
 * Skipping exact zeros / padded rows is exact.  write_pad != 0 also writes the constant rows
 * act(bias) of padded nodes (needed when the full [B,N,H] tensor is read afterwards).
 * Requirements of the fused kernel: N <= 128, Din % 32 == 0, K % 4 == 0, K <= 32, H % 4 == 0,
 * H <= 128, E1 <= 16; W is [H, (S+E1)*Din].  Returns LNB_ERR_UNSUPPORTED otherwise (callers
 * use the unfused ops).
 * ------------------------------------------------------------------------------------- */
int lnb_graph_prepare(lnb_stream_t stream, const float* L, const float* Q, int B, int N, int E1,
                      int K, float* ell_val, uint8_t* ell_idx, int32_t* ell_max, int32_t* gext,
                      int32_t* tiles, int32_t* rowmap, int32_t* nrows, int flags);
int lnb_spectral_conv_fused(lnb_stream_t stream, const float* X, const float* Q, const float* coeff,
                            const float* ell_val, const uint8_t* ell_idx, const int32_t* ell_max,
                            const int32_t* gext, const int32_t* tiles, const float* W_hi,
                            const float* W_lo, const float* bias, int B, int N, int Din, int E1,
                            int K, int S, int H, int relu, int write_pad, float* out);

/* ---------------------------------------------------------------------------------------
 * GPU-side batch construction from SPARSE per-molecule records (replaces, on the device, the host
 * pipeline utils/data_helper.py:92-116,155-156 (L4 = D^-1/2 (A + I) D^-1/2 of every bond channel and
 * of the simple graph) + dataset/qm8.py:57-90,220-291 (zero padding / stacking of node_feat,
 * node_mask, L, (D, V)) and the dense pass of lnb_graph_prepare).  Inputs, all device pointers:
 *   sizes [B] real nodes per graph; node_ptr [B+1] their prefix sums; node_feat [node_ptr[B]] atom
 *   ids of the real nodes; edge_ptr [B+1]; edges [edge_ptr[B]][4] bytes {u, v, bond type, 0}
 *   (undirected bonds listed once, local node indices); V_rows [node_ptr[B], K] Ritz vectors of the
 *   real nodes; inv_sqrt_deg [256] fp64 table of deg^-1/2 (entry 0 = 0) from the host's numpy, so
 *   the fp64 products (scale_i * m_ij) * scale_j and their single rounding to fp32 are bit-identical
 *   to the reference's preprocessing.
 * Outputs: everything lnb_graph_prepare emits (same layouts, same bits: ell_val / ell_idx / ell_max /
 * gext / tiles / rowmap / nrows), the padded node_ids [B,N] int64, mask [B,N] uint8 and
 * V [B,N,K] that lnb_spectral_stack_forward reads, and -- only when L_dense != NULL -- the padded dense
 * operators [B,N,N,E1] exactly as the reference's collate builds them.  flags as lnb_graph_prepare.
 * Limits: N <= 128, 2 <= E1 <= 16, degrees < 255.
 * ------------------------------------------------------------------------------------- */
int lnb_graph_prepare_sparse(lnb_stream_t stream, const int32_t* sizes, const int32_t* node_ptr,
                             const int32_t* node_feat, const int32_t* edge_ptr, const uint8_t* edges,
                             const float* V_rows, const double* inv_sqrt_deg, int B, int N, int E1,
                             int K, int flags, float* ell_val, uint8_t* ell_idx, int32_t* ell_max,
                             int32_t* gext, int32_t* tiles, int32_t* rowmap, int32_t* nrows,
                             int64_t* node_ids, uint8_t* mask, float* V, float* L_dense);

/* Packed variant: the whole sparse batch in ONE contiguous, 16-byte aligned device buffer, so a step
 * costs a single H2D copy of exactly the bytes present (eight ranks issuing seven small copies each
 * were host-bound).  Layout, all offsets in bytes and multiples of 16, int32 header first:
 *   hdr[0] = 0x4c4e4231 ("LNB1"), hdr[1] = B, hdr[2] = K, hdr[3] = off(sizes [B] i32),
 *   hdr[4] = off(node_ptr [B+1] i32), hdr[5] = off(edge_ptr [B+1] i32), hdr[6] = off(D [B,K] f32),
 *   hdr[7] = off(node_feat [sum n] i32), hdr[8] = off(V_rows [sum n, K] f32),
 *   hdr[9] = off(edges [sum E][4] u8), hdr[10] = total bytes, hdr[11] = off(tiles [B+2] i32),
 *   hdr[12] = off(krow_ptr [B+1] i32) (both 0 when absent); hdr[3..6], hdr[11], hdr[12] depend on (B, K)
 *   only, so D and the tile table sit at fixed addresses of a reused buffer (lnb_ritz_power_table and
 *   lnb_spectral_stack_forward read them there).
 * The kernel derives its input pointers from the header on the device.  flags bit 1
 * (LNB_PACKED_HOST_TILES): the host knows every graph's extents, so it ships the packed-tile table
 * (same next-fit rule as lnb_graph_prepare: consecutive graphs, sum n <= 128, sum ceil4(k_eff) <= 128,
 * <= 32 graphs) and the prefix sums krow_ptr of k_eff; the kernel expands the Ritz row list itself and
 * NO tile-assignment launch follows (the `tiles` argument is then unused: pass the blob's segment to
 * the stack kernel). */
#define LNB_PACKED_HOST_TILES 2
int lnb_graph_prepare_sparse_packed(lnb_stream_t stream, const uint8_t* blob, const double* inv_sqrt_deg,
                                    int B, int N, int E1, int K, int flags, float* ell_val,
                                    uint8_t* ell_idx, int32_t* ell_max, int32_t* gext, int32_t* tiles,
                                    int32_t* rowmap, int32_t* nrows, int64_t* node_ids, uint8_t* mask,
                                    float* V, float* L_dense);

/* ---------------------------------------------------------------------------------------
 * The whole convolution stack (and optionally the embedding gather in front and the readout
 * behind it) in ONE persistent kernel: every CTA keeps its packed tile's state in shared memory
 * across layers, so between layers nothing touches HBM.  Layer l uses rows [l*H, (l+1)*H) of
 * the stacked split weights W_hi / W_lo [num_layers*H, Kw] (columns beyond (S+E1)*Din[l] zero),
 * bias + l*H and coeff + l*coeff_layer_stride.  Input: X [B,N,Din[0]] or node_ids [B,N] +
 * emb_table [emb_rows, Din[0]] (model/lanczos_net.py:154).  Outputs: out_state [B,N,H] (may be
 * NULL) and / or score [B,P] from the fused readout (model/lanczos_net.py:185-194; mask may be
 * NULL = mean over all N nodes; P <= 48).  Same shape limits as lnb_spectral_conv_fused, plus
 * Din[l>0] == H, num_layers <= 8 and a 16-byte aligned bias.
 * ------------------------------------------------------------------------------------- */
typedef struct lnb_spectral_stack {
  const float* X; const int64_t* node_ids; const float* emb_table;
  const float* Q; const float* coeff; int64_t coeff_layer_stride;
  const float* ell_val; const uint8_t* ell_idx; const int32_t* ell_max; const int32_t* gext;
  const int32_t* tiles;
  const float* W_hi; const float* W_lo; const float* bias;
  float* out_state;
  const float* W_out; const float* b_out; const float* w_att; const float* b_att;
  const uint8_t* mask; float* score;
  int32_t Din[8];
  int32_t num_layers, Kw, emb_rows, P, write_pad;
  int32_t B, N, E1, K, S, H, relu;
} lnb_spectral_stack;
int lnb_spectral_stack_forward(lnb_stream_t stream, const lnb_spectral_stack* desc /* host */);

/* ---------------------------------------------------------------------------------------
 * Embedding rows (model/lanczos_net.py:154): out[r, :] = table[idx[r], :].
 * ------------------------------------------------------------------------------------- */
int lnb_embedding_rows(lnb_stream_t stream, const int64_t* idx, const float* table,
                       int64_t rows, int num_embeddings, int dim, float* out);

/* ---------------------------------------------------------------------------------------
 * Ritz-value power table (model/lanczos_net.py:146-149): table[b,k,s] = D[b,k] ** powers[s],
 * correctly rounded from a double-precision pow.  The per-layer filter MLP
 * (model/lanczos_net.py:109-113) is then four lnb_batched_gemm / lnb_linear_tf32x3 calls over
 * the B*K rows, batched over all layers at once because the input does not depend on the
 * layer state.  powers: host pointer to S ints (S <= 32).
 * ------------------------------------------------------------------------------------- */
int lnb_ritz_power_table(lnb_stream_t stream, const float* D, int64_t rows, const int* powers,
                         int S, float* table /* [rows, S] */);

/* ---------------------------------------------------------------------------------------
 * Ritz-value filter MLPs of all layers in one persistent tcgen05 kernel
 * (model/lanczos_net.py:47-58,109-113): coeff[l, r, :] = MLP_l(table[r, :]) for the rows r listed
 * in rowmap (nrows[0] entries; both NULL = all Rall rows).  The four Linear stages of a
 * (row tile, layer) item run back to back with the 128 x hidden activations kept in shared
 * memory.  W_hi/W_lo: tf32 split of the stacked weights [L*(3*hidden+S), hidden]: per layer the
 * rows of stage 0 (input columns zero-padded from S to hidden), stage 1, stage 2, stage 3
 * (S rows); bias_all uses the same row indexing.  lnb_ritz_rowmap builds the compact row list
 * {b*K + k : k < k_eff(b)} from the extents of lnb_graph_prepare (rows of zero-padded Ritz pairs
 * multiply zero Ritz vectors downstream and are skipped; their coeff entries stay unwritten).
 * Requirements: S <= 32, hidden % 32 == 0, hidden <= 128 (else LNB_ERR_UNSUPPORTED).
 * ------------------------------------------------------------------------------------- */
int lnb_ritz_rowmap(lnb_stream_t stream, const int32_t* gext, int B, int K, int32_t* rowmap,
                    int32_t* nrows);
int lnb_ritz_filter_mlp(lnb_stream_t stream, const float* table, const int32_t* rowmap,
                        const int32_t* nrows, const float* W_hi, const float* W_lo,
                        const float* bias_all, int Rall, int L, int S, int Hd, float* coeff);

/* ---------------------------------------------------------------------------------------
 * Readout (model/lanczos_net.py:185-194, ada_lanczos_net.py:350-361):
 *   y[b,n,:] = (W_out state[b,n,:] + b_out) * sigmoid(w_att . state[b,n,:] + b_att)
 *   score[b,:] = mean over n with mask[b,n] != 0 (mask == NULL -> all n)
 * ------------------------------------------------------------------------------------- */
int lnb_readout(lnb_stream_t stream, const float* state, const float* W_out, const float* b_out,
                const float* w_att, const float* b_att, const uint8_t* mask, int B, int N, int H,
                int P, float* score /* [B,P] */);

/* ---------------------------------------------------------------------------------------
 * Operator chain on channel 0 of L [B,N,N,E1], per graph, starting from X [B,N,D]:
 *   chebyshev == 0: w_s = L_0 w_{s-1} (w_0 = X), s = 1..steps   (model/dcnn.py:88-92, the short
 *                   diffusion walk of model/lanczos_net.py:164-169);
 *   chebyshev != 0: s_0 = L_0 X, s_k = 2 L_0 s_{k-1} - s_{k-2} with s_{-1} = X, k < steps
 *                   (model/cheby_net.py:88-93).
 * Result number i (0-based) is written to out[b, n, (out_col0 + block_of_step[i]) * D + d] when
 * block_of_step[i] >= 0 (host array of `steps` ints); out strides in elements.  One launch for
 * the whole chain: operator and walk stay in shared memory / registers.  N <= 32, steps <= 64
 * (LNB_ERR_UNSUPPORTED otherwise: callers use lnb_batched_gemm per step).
 * ------------------------------------------------------------------------------------- */
int lnb_operator_chain(lnb_stream_t stream, const float* L, const float* X, int B, int N, int E1,
                       int D, int steps, int chebyshev, const int* block_of_step /* host */,
                       float* out, int64_t out_batch_stride, int64_t out_row_stride, int out_col0);

/* ---------------------------------------------------------------------------------------
 * The whole message matrix of a general-shape spectral convolution layer in one launch
 * (model/lanczos_net.py:157-180, model/ada_lanczos_net.py:321-345):
 *   out[b, n, :] = [ (L_0^k X)[n] : k selected ] ++ [ (Q G_s Q^T X)[n] : s < S ] ++ [ (L_e X)[n] : e < E1 ]
 * L [B,N,N,E1], X [B,N,D], Q [B,N,K]; filt = G [B,S,K,K] symmetric blocks when dense_filter != 0
 * (AdaLanczosNet's learned filter), else the diagonal coefficients [B,K,S] (LanczosNet).  Short walk:
 * step i (1-based) goes to column block block_of_step[i-1] (< 0: not stored; host array of
 * short_steps ints), the long scales to blocks n_short + s, the edge types to n_short + S + e; every
 * block is D columns wide; out strides in elements.  One CTA per graph, thread per feature column,
 * operators and filters in shared memory, every intermediate in registers.  N <= 32, K <= 32,
 * E1 <= 16, S <= 8 (LNB_ERR_UNSUPPORTED otherwise: callers compose lnb_batched_gemm calls).
 * ------------------------------------------------------------------------------------- */
int lnb_graph_messages(lnb_stream_t stream, const float* L, const float* X, const float* Q,
                       const float* filt, int B, int N, int E1, int D, int K, int S, int dense_filter,
                       int short_steps, const int* block_of_step /* host */, int n_short, float* out,
                       int64_t out_batch_stride, int64_t out_row_stride);

/* ---------------------------------------------------------------------------------------
 * Gaussian-kernel graph Laplacian (model/ada_lanczos_net.py:101-137, adjacency from :310-311):
 *   adj = (L[b,i,j,0] != 0);  dist2 = |x_i - x_j|^2;  sigma2 = mean over all N^2 pairs;
 *   A = exp(-dist2/sigma2) * adj;  d = (rowsum + [rowsum==0])^-1/2;  out = d_i A_ij d_j
 * L has E1 channels innermost (dataset/qm8.py:262); only channel 0 is read.
 * ------------------------------------------------------------------------------------- */
int lnb_gaussian_laplacian(lnb_stream_t stream, const float* x, const float* L, int B, int N,
                           int Dx, int E1, float* out /* [B,N,N] */);

/* ---------------------------------------------------------------------------------------
 * Batched K-step Lanczos tridiagonalisation with full double re-orthogonalisation and the
 * reference's masking rules (model/ada_lanczos_net.py:139-247).  q1 is the raw start vector
 * (the randn draw of :161); mask (uint8, may be NULL) zeroes padded nodes.
 * Outputs: T [B,K,K] dense tridiagonal, Q [B,N,K], alpha [B,K], beta [B,K] (beta[b,K-1]=0),
 * idx [B] int32 = number of retained Krylov directions (:208-211).
 * ------------------------------------------------------------------------------------- */
int lnb_lanczos_tridiag(lnb_stream_t stream, const float* A, const uint8_t* mask, const float* q1,
                        int B, int N, int K, float* T, float* Q, float* alpha, float* beta,
                        int32_t* idx);

/* ---------------------------------------------------------------------------------------
 * Ritz pairs of the Lanczos tridiagonal: implicit-shift QL on (alpha, beta) with the
 * rotations applied to Q, so ritz_vec = Q S directly.  Ordered by descending |theta|
 * (the reference's Ritz ordering, utils/data_helper.py:217-223); ties keep ascending index.
 * status[b] = 0 ok, >0 = QL sweeps exhausted on that graph.
 * ------------------------------------------------------------------------------------- */
int lnb_tridiag_ritz(lnb_stream_t stream, const float* alpha, const float* beta, const float* Q,
                     int B, int N, int K, float* theta /* [B,K] */, float* ritz_vec /* [B,N,K] */,
                     int32_t* status /* [B] */);

/* ---------------------------------------------------------------------------------------
 * The north-star pipeline in ONE launch: operator -> K-step Lanczos (rules of
 * model/ada_lanczos_net.py:139-247, as lnb_lanczos_tridiag) -> implicit-shift QL on (alpha, beta)
 * -> Ritz vectors V = Q S ordered by descending |theta| (utils/data_helper.py:217-223) -- the pair
 * (theta, V) is what utils/data_helper.py:169-226 + dataset/qm8.py:265-291 hand to
 * LanczosNet.forward as (D, V).  One group of 32..512 threads per graph; the dense padded operator
 * A [B,N,N] is read from HBM exactly once (its non-zeros are packed into shared memory, exact zeros
 * contribute nothing to A q); the Krylov basis, (alpha, beta) and the QL rotations never leave the
 * SM.  T, Q may be NULL (not written).  theta / ritz_vec / status may be NULL together: then only
 * the tridiagonalisation is produced (AdaLanczosNet).  status[b]: bit 0 = QL sweeps exhausted,
 * bit 1 = the graph's non-zeros did not fit on chip and its rows were streamed per iteration.
 * flags: 0 = the reference's masking rules (idx = min(#valid betas, #real nodes) directions and node
 * rows kept; the alpha of the breakdown step dropped, ada_lanczos_net.py:207-237);
 * LNB_LANCZOS_PROPER = the textbook Krylov factorisation (m = #valid + 1 vectors, T_m with m alphas
 * and m-1 betas, no row masking; idx[b] = m) whose Ritz values are eigenvalues of A -- the mode of
 * the online (D, V) provider.
 * Limits: N <= 1024, K <= 64 and a basis of K*(N+1) floats within shared memory
 * (LNB_ERR_UNSUPPORTED otherwise: use lnb_lanczos_tridiag + lnb_tridiag_ritz).
 * ------------------------------------------------------------------------------------- */
#define LNB_LANCZOS_PROPER 1
int lnb_lanczos_ritz(lnb_stream_t stream, const float* A, const uint8_t* mask, const float* q1,
                     int B, int N, int K, int flags, float* T, float* Q, float* alpha, float* beta,
                     int32_t* idx, float* theta /* [B,K] */, float* ritz_vec /* [B,N,K] */,
                     int32_t* status /* [B] */);

/* ---------------------------------------------------------------------------------------
 * Powers of the tridiagonal for the learned filter (model/ada_lanczos_net.py:262-274):
 *   out[b, r, s, c] = (T_b ** powers[s])[r, c]    (the MLP input layout r*S*K + s*K + c)
 * powers: host pointer to S strictly increasing positive ints.
 * ------------------------------------------------------------------------------------- */
int lnb_tridiag_powers(lnb_stream_t stream, const float* T, int B, int K, const int* powers, int S,
                       float* out /* [B,K,S,K] */);

/* Symmetrised filter blocks (model/ada_lanczos_net.py:275-278):
 *   G[b,s,r,c] = 0.5 * (Y[b, r*K*S + c*S + s] + Y[b, c*K*S + r*S + s])                     */
int lnb_symmetrize_filters(lnb_stream_t stream, const float* Y, int B, int K, int S,
                           float* G /* [B,S,K,K] */);

/* Profiling aid (not used by the product path): register a device buffer of (number of SMs) x 32 uint64 that
 * the tcgen05 kernels fill with per-CTA clock64 totals per phase; NULL disables. */
int lnb_debug_set_prof(unsigned long long* buf);

#ifdef __cplusplus
}
#endif
#endif /* LANCZOSNET_B200_H_ */
