"""Training path: the forward AND backward of the spectral-convolution models as
``torch.autograd.Function``s whose arithmetic runs in this library's CUDA kernels.

The reference trains through autograd over ``torch.bmm`` / ``nn.Linear``
(runner/qm8_runner.py:188-259, ``train_loss.backward()`` at :247).  Here every contraction of the
forward and of its adjoint goes through the C ABI:

  * dense layers ``act(x W^T + b)`` and their two adjoint products (``g W`` and ``g^T x``): the
    tcgen05 3xTF32 kernel (lnb_linear_tf32x3) -- the big Linear of the graph-conv layer is 97 % of
    the flops of a training step -- or, for layers too small to amortise that kernel's set-up, the
    strided fp32 GEMM reading W / W^T / g^T in place;
  * operator products ``L_e X``, ``L_e^T G`` (channel-innermost operators read in place through
    their strides), ``V^T X``, ``V diag(f) U`` and their adjoints: the strided batched GEMM
    (lnb_batched_gemm);
  * the embedding gradient: the scatter-add of the reference's own ``unsorted_segment_sum`` op
    (lnb_unsorted_segment_sum_forward).

PyTorch supplies what it supplies everywhere in this package -- tensor memory, streams, the autograd
tape -- plus the pointwise glue of the adjoint (ReLU masks, sigmoid gate, masked mean, the loss).
The training forward is the UNFUSED formulation (activations have to exist to be differentiated);
the one-launch fused stack remains the inference path.  Gradients are checked against
``torch.autograd`` over the fp64 CPU oracle in tests/test_gpu_train.py.
"""
import torch

from . import ops

__all__ = ['dense', 'operator_messages', 'spectral_messages', 'embedding', 'ritz_stack_train',
           'dcnn_train', 'cheby_train', 'gated_readout', 'bmm', 'ada_train', 'GraphedStep']


def _pad_cols(x, mult=4):
  k = x.shape[1]
  if k % mult == 0:
    return x.contiguous()
  return torch.nn.functional.pad(x, (0, mult - k % mult)).contiguous()


def _matmul_nt(a, w):
  """a [M,K] @ w[N,K]^T on the tcgen05 3xTF32 kernel; K is zero padded to a multiple of 4."""
  a, w = _pad_cols(a.float()), _pad_cols(w.float())
  w_hi, w_lo = ops.split_tf32(w)
  return ops.linear_tf32x3(a, w_hi, w_lo, None, False)


# Below this many flops (2 M N K) a dense layer is launch bound on the persistent tcgen05 kernel (TMEM
# allocation, barrier set-up, operand splits, padded / transposed copies for its two adjoint products):
# the strided fp32 GEMM reads W, W^T, g^T in place and is exact fp32.  At the reference's batch size (64
# molecules) this is every layer but the graph-conv Linear, which keeps 95 % of the flops on the tensor cores.
_SMALL_DENSE_FLOPS = 1.5e8


def _small_gemm(A, a_str, Bm, b_str, M, N, K, bias=None, relu=False):
  C = torch.empty((M, N), device=A.device, dtype=torch.float32)
  ops.bgemm(A, (0, 0) + a_str, Bm, (0, 0) + b_str, C, (0, 0, N, 1), 1, 1, M, N, K, bias=bias, relu=relu)
  return C


class _Dense(torch.autograd.Function):
  """y = act(x W^T + b): nn.Linear (+ ReLU) of model/lanczos_net.py:109-113,180-181,188-189."""

  @staticmethod
  def forward(ctx, x, weight, bias, relu):
    M, K = x.shape
    N = weight.shape[0]
    ctx.small = 2.0 * M * N * K < _SMALL_DENSE_FLOPS
    if ctx.small:
      x, weight = x.float().contiguous(), weight.float().contiguous()
      y = _small_gemm(x, (K, 1), weight, (1, K), M, N, K, bias.float().contiguous() if bias is not None else None, relu)
    else:
      xp, wp = _pad_cols(x.float()), _pad_cols(weight.float())
      w_hi, w_lo = ops.split_tf32(wp)
      y = ops.linear_tf32x3(xp, w_hi, w_lo, bias, relu)
    ctx.relu = bool(relu)
    ctx.save_for_backward(x, weight, y if relu else None)
    ctx.has_bias = bias is not None
    return y

  @staticmethod
  def backward(ctx, gy):
    x, weight, y = ctx.saved_tensors
    gy = gy.contiguous()
    if ctx.relu:
      gy = gy * (y > 0).to(gy.dtype)
    gx = gw = gb = None
    M, K = x.shape
    N = weight.shape[0]
    if ctx.needs_input_grad[0]:                                     # g W
      gx = (_small_gemm(gy, (N, 1), weight, (K, 1), M, K, N) if ctx.small
            else _matmul_nt(gy, weight.t())[:, :K])
    if ctx.needs_input_grad[1]:                                     # g^T x
      # few output tiles, long contraction over the rows: the tensor-core kernel wins at every size
      gw = _matmul_nt(gy.t(), x.t())[:, :K]
    if ctx.has_bias and ctx.needs_input_grad[2]:
      gb = gy.sum(dim=0)
    return gx, gw, gb, None


def dense(x, weight, bias, relu=False):
  return _Dense.apply(x, weight, bias, relu)


class _OperatorMessages(torch.autograd.Function):
  """msg[b, n, e*D:(e+1)*D] = (L[b, :, :, e] X[b])[n]   for the channels e in ``channels``
  (model/lanczos_net.py:177-178); adjoint: gX = sum_e L_e^T g_e."""

  @staticmethod
  def forward(ctx, L, X, c0, nc):
    B, N, D = X.shape
    E1 = L.shape[3]
    X = X.contiguous()
    msg = torch.empty((B, N, nc * D), device=X.device, dtype=torch.float32)
    ops.bgemm(L, (N * N * E1, 1, N * E1, E1), X, (N * D, 0, D, 1), msg, (N * nc * D, D, nc * D, 1),
              B, nc, N, D, N, a_off=c0)
    ctx.save_for_backward(L)
    ctx.c0, ctx.nc = c0, nc
    return msg

  @staticmethod
  def backward(ctx, g):
    (L,) = ctx.saved_tensors
    B, N, E1 = L.shape[0], L.shape[1], L.shape[3]
    nc = ctx.nc
    D = g.shape[2] // nc
    g = g.contiguous()
    tmp = torch.empty((B, nc, N, D), device=g.device, dtype=torch.float32)
    # A[m = column c][k = row r] = L[b, r, c, e]: the transpose through swapped strides
    ops.bgemm(L, (N * N * E1, 1, E1, N * E1), g, (N * nc * D, D, nc * D, 1), tmp,
              (nc * N * D, N * D, D, 1), B, nc, N, D, N, a_off=ctx.c0)
    return None, tmp.sum(dim=1), None, None


def operator_messages(L, X, c0=0, nc=None):
  return _OperatorMessages.apply(L, X, c0, L.shape[3] - c0 if nc is None else nc)


class _SpectralMessages(torch.autograd.Function):
  """msg[b, n, s*D:(s+1)*D] = (V diag(F[:, :, s]) V^T X)[b, n]  in factored form
  (model/lanczos_net.py:114-123,172-175); F = filter coefficients [B,K,S] (differentiable: they are
  the output of the learned spectral filter), V is data."""

  @staticmethod
  def forward(ctx, V, X, F):
    B, N, D = X.shape
    K, S = V.shape[2], F.shape[2]
    X, F = X.contiguous(), F.contiguous()
    U = torch.empty((B, K, D), device=X.device, dtype=torch.float32)
    ops.bgemm(V, (N * K, 0, 1, K), X, (N * D, 0, D, 1), U, (K * D, 0, D, 1), B, 1, K, D, N)
    msg = torch.empty((B, N, S * D), device=X.device, dtype=torch.float32)
    ops.bgemm(V, (N * K, 0, K, 1), U, (K * D, 0, D, 1), msg, (N * S * D, D, S * D, 1), B, S, N, D, K,
              kscale=F, s_str=(K * S, 1, S))
    ctx.save_for_backward(V, U, F)
    return msg

  @staticmethod
  def backward(ctx, g):
    V, U, F = ctx.saved_tensors
    B, N, K = V.shape
    S = F.shape[2]
    D = U.shape[2]
    g = g.contiguous()
    T = torch.empty((B, S, K, D), device=g.device, dtype=torch.float32)        # T_s = V^T g_s
    ops.bgemm(V, (N * K, 0, 1, K), g, (N * S * D, D, S * D, 1), T, (S * K * D, K * D, D, 1),
              B, S, K, D, N)
    gF = (T * U.unsqueeze(1)).sum(dim=3).permute(0, 2, 1).contiguous()         # [B,K,S]
    gU = (T * F.permute(0, 2, 1).unsqueeze(3)).sum(dim=1).contiguous()         # [B,K,D]
    gX = torch.empty((B, N, D), device=g.device, dtype=torch.float32)
    ops.bgemm(V, (N * K, 0, K, 1), gU, (K * D, 0, D, 1), gX, (N * D, 0, D, 1), B, 1, N, D, K)
    return None, gX, gF


def spectral_messages(V, X, F):
  return _SpectralMessages.apply(V, X, F)


class _Embedding(torch.autograd.Function):
  """state = table[ids] (model/lanczos_net.py:154); the gradient of the table is the scatter-add of
  the reference's own unsorted_segment_sum op."""

  @staticmethod
  def forward(ctx, ids, table):
    ctx.save_for_backward(ids)
    ctx.rows = table.shape[0]
    return ops.embedding_rows(ids, table)

  @staticmethod
  def backward(ctx, g):
    (ids,) = ctx.saved_tensors
    D = g.shape[-1]
    flat = g.reshape(1, -1, D).contiguous()
    seg = ids.reshape(1, -1).clamp(0, ctx.rows - 1)
    return None, ops.segment_sum_forward(flat, seg, ctx.rows)[0]


def embedding(ids, table):
  return _Embedding.apply(ids.long(), table)


def ritz_stack_train(model, state, node_ids, L, D, V, mask):
  """Differentiable convolution stack + readout of LanczosNet / LanczosNetGeneral / GCN
  (model/lanczos_net.py:125-199): same math, same parameter tensors as the inference path."""
  L = L.float().contiguous()
  if node_ids is not None:
    state = embedding(node_ids, model.embedding.weight)
  else:
    state = state.float().contiguous()
  B, N = state.shape[0], state.shape[1]
  S = model.num_scale_long
  short = list(model.short_diffusion_dist)
  table = None
  if S > 0:
    V = V.float().contiguous()
    table = ops.ritz_power_table(D.float().contiguous(), model.long_diffusion_dist)   # [B,K,S], data
    K = table.shape[1]
  for t in range(model.num_layer):
    msgs = []
    if short:                                   # walk <- L0 walk (lanczos_net.py:164-169)
      walk = state
      for step in range(1, max(short) + 1):
        walk = operator_messages(L, walk, 0, 1)
        if step in short:
          msgs.append(walk)
    if S > 0:
      if model.spectral_filter_kind == 'MLP':
        seq = model.spectral_filter[t]
        h = table.reshape(B * K, S)
        for i in (0, 2, 4, 6):
          h = dense(h, seq[i].weight, seq[i].bias, i != 6)
        F = h.reshape(B, K, S)
      else:
        F = table
      msgs.append(spectral_messages(V, state, F))
    msgs.append(operator_messages(L, state))
    msg = torch.cat(msgs, dim=2) if len(msgs) > 1 else msgs[0]
    lin = model.filter[t]
    state = dense(msg.reshape(B * N, -1), lin.weight, lin.bias, True).reshape(B, N, -1)
    if model.training and model.dropout > 0.0:
      state = torch.nn.functional.dropout(state, model.dropout, True)
  return gated_readout(model, state, mask)


def gated_readout(model, state, mask):
  """Gated masked-mean readout shared by all models (lanczos_net.py:185-194): the two Linears in the
  library's dense kernel, the pointwise gate / mean on the autograd tape."""
  B, N = state.shape[0], state.shape[1]
  head, att = model.filter[model.num_layer], model.att_func[0]
  flat = state.reshape(B * N, -1)
  y = dense(flat, head.weight, head.bias, False).reshape(B, N, -1)
  gate = torch.sigmoid(dense(flat, att.weight, att.bias, False)).reshape(B, N, 1)
  y = y * gate
  if mask is None:
    return y.mean(dim=1)
  m = (mask != 0).to(y.dtype).unsqueeze(2)
  return (y * m).sum(dim=1) / m.sum(dim=1)


def dcnn_train(model, node_ids, L, mask):
  """Differentiable DCNN (model/dcnn.py:64-124): per layer the edge-type products, then the walk
  L_0^k X for k in diffusion_dist, concatenated EDGES FIRST (:98), Linear + ReLU."""
  L = L.float().contiguous()
  state = embedding(node_ids, model.embedding.weight)
  B, N = state.shape[0], state.shape[1]
  dist = set(model.diffusion_dist)
  for t in range(model.num_layer):
    msgs = [operator_messages(L, state)]
    walk = state
    for step in range(1, model.max_dist + 1):
      walk = operator_messages(L, walk, 0, 1)
      if step in dist:
        msgs.append(walk)
    lin = model.filter[t]
    state = dense(torch.cat(msgs, dim=2).reshape(B * N, -1), lin.weight, lin.bias, True).reshape(B, N, -1)
    if model.training and model.dropout > 0.0:
      state = torch.nn.functional.dropout(state, model.dropout, True)
  return gated_readout(model, state, mask)


def cheby_train(model, node_ids, L, mask):
  """Differentiable ChebyNet (model/cheby_net.py:64-124): s_0 = L_0 X, s_k = 2 L_0 s_{k-1} - s_{k-2}
  with s_{-1} = X (the reference's index -1 is its LAST slot, :88-93), bond-type products for e >= 1,
  cat(edges + [s_0 .. s_{order-1}] + [X]) (:99), Linear + ReLU."""
  L = L.float().contiguous()
  state = embedding(node_ids, model.embedding.weight)
  B, N, E1 = state.shape[0], state.shape[1], L.shape[3]
  order = model.polynomial_order
  for t in range(model.num_layer):
    scale = [None] * (order + 1)
    scale[-1] = state
    scale[0] = operator_messages(L, state, 0, 1)
    for kk in range(1, order):
      scale[kk] = 2.0 * operator_messages(L, scale[kk - 1], 0, 1) - scale[kk - 2]
    msgs = ([operator_messages(L, state, 1, E1 - 1)] if E1 > 1 else []) + scale
    lin = model.filter[t]
    state = dense(torch.cat(msgs, dim=2).reshape(B * N, -1), lin.weight, lin.bias, True).reshape(B, N, -1)
    if model.training and model.dropout > 0.0:
      state = torch.nn.functional.dropout(state, model.dropout, True)
  return gated_readout(model, state, mask)


class _BMM(torch.autograd.Function):
  """C[b] = A[b] @ Bm[b] on the strided batched GEMM, differentiable in both operands
  (gA = g Bm^T, gB = A^T g through swapped strides -- no transposed copies)."""

  @staticmethod
  def forward(ctx, A, Bm):
    A, Bm = A.contiguous(), Bm.contiguous()
    nb, M, K = A.shape
    N = Bm.shape[2]
    C = torch.empty((nb, M, N), device=A.device, dtype=torch.float32)
    ops.bgemm(A, (M * K, 0, K, 1), Bm, (K * N, 0, N, 1), C, (M * N, 0, N, 1), nb, 1, M, N, K)
    ctx.save_for_backward(A, Bm)
    return C

  @staticmethod
  def backward(ctx, g):
    A, Bm = ctx.saved_tensors
    nb, M, K = A.shape
    N = Bm.shape[2]
    g = g.contiguous()
    gA = gB = None
    if ctx.needs_input_grad[0]:            # [M,N] @ [N,K]: B operand = Bm^T via (k stride 1, n stride N)
      gA = torch.empty_like(A)
      ops.bgemm(g, (M * N, 0, N, 1), Bm, (K * N, 0, 1, N), gA, (M * K, 0, K, 1), nb, 1, M, K, N)
    if ctx.needs_input_grad[1]:            # [K,M] @ [M,N]: A operand = A^T via (m stride 1, k stride K)
      gB = torch.empty_like(Bm)
      ops.bgemm(A, (M * K, 0, 1, K), g, (M * N, 0, N, 1), gB, (K * N, 0, N, 1), nb, 1, K, N, M)
    return gA, gB


def bmm(A, Bm):
  return _BMM.apply(A.float(), Bm.float())


_EPS = 1.1920928955078125e-07       # np.finfo(np.float32).eps (ada_lanczos_net.py:8)


def _gaussian_laplacian_train(x, adj):
  """Learned operator of model/ada_lanczos_net.py:101-137 on the autograd tape: Gaussian kernel of the
  embedding distances (sigma^2 = mean over all N^2 pairs, padded ones included), masked by the
  adjacency, symmetrically normalised."""
  diff = x.unsqueeze(1) - x.unsqueeze(2)
  dist2 = (diff * diff).sum(dim=3)
  sigma2 = dist2.reshape(dist2.shape[0], -1).mean(dim=1).reshape(-1, 1, 1)
  A = torch.exp(-dist2 / sigma2) * adj
  rs = A.sum(dim=2, keepdim=True)
  d = (rs + (rs == 0).to(A.dtype)).pow(-0.5)
  return d * A * d.transpose(1, 2)


def _lanczos_train(A, mask, q1, K):
  """Differentiable K-step Lanczos with the reference's rules (model/ada_lanczos_net.py:139-247);
  operator products through ``bmm``, re-orthogonalisation as two block Gram-Schmidt passes (the
  formulation of the inference kernel), the acceptance / masking logic as data (no gradient)."""
  B, N = A.shape[0], A.shape[1]
  iters = min(N, K)
  q = q1.reshape(B, N, 1).to(A.dtype)
  nreal = torch.full((B,), N, device=A.device, dtype=torch.long)
  if mask is not None:
    fm = (mask != 0).reshape(B, N, 1).to(A.dtype)
    q = q * fm
    nreal = fm.sum(dim=1).reshape(B).long()
  q = q / q.norm(dim=1, keepdim=True)
  basis, alphas, betas, valids = [q], [], [], []
  prev, beta_prev = torch.zeros_like(q), torch.zeros((B, 1, 1), device=A.device, dtype=A.dtype)
  ok = torch.ones((B, 1, 1), device=A.device, dtype=A.dtype)
  for i in range(iters):
    cur = basis[i]
    z = bmm(A, cur)
    a = (cur * z).sum(dim=1, keepdim=True)
    z = z - a * cur - beta_prev * prev
    if i > 0:
      Qb = torch.cat(basis[:i], dim=2)                               # [B,N,i]
      scale = 1.0 / ((Qb * Qb).sum(dim=1, keepdim=True) + _EPS)       # [B,1,i]
      for _ in range(2):
        c = bmm(Qb.transpose(1, 2), z) * scale.transpose(1, 2)       # [B,i,1]
        z = z - bmm(Qb, c)
    b = z.norm(dim=1, keepdim=True)
    ok = ok * (b.detach() >= 1.0e-4).to(A.dtype)
    valids.append(ok)
    alphas.append(a)
    betas.append(b)
    basis.append(z * ok / (b + _EPS))
    prev, beta_prev = cur, b
  alpha = torch.cat(alphas, dim=1).squeeze(2)
  valid = torch.cat(valids, dim=1).squeeze(2)
  idx = torch.minimum(valid.sum(dim=1).long(), nreal)
  col = torch.arange(iters, device=A.device).unsqueeze(0)
  valid = valid * (col < idx.unsqueeze(1)).to(A.dtype)
  alpha = alpha * valid
  T = torch.diag_embed(alpha)
  if iters > 1:
    beta = torch.cat(betas[:-1], dim=1).squeeze(2) * valid[:, :-1]
    T = T + torch.diag_embed(beta, offset=1) + torch.diag_embed(beta, offset=-1)
  Q = torch.cat(basis[:iters], dim=2)
  row = torch.arange(N, device=A.device).reshape(1, N, 1)
  Q = Q * (valid.unsqueeze(1) * (row < idx.reshape(B, 1, 1)).to(A.dtype))
  if iters < K:
    T = torch.nn.functional.pad(T, (0, K - iters, 0, K - iters))
    Q = torch.nn.functional.pad(Q, (0, K - iters))
  return T, Q


def ada_train(model, node_ids, L, mask, q1):
  """Differentiable AdaLanczosNet (model/ada_lanczos_net.py:288-368): embedding -> learned Gaussian
  Laplacian -> Lanczos -> learned filter on the powers of T (the 4096-wide MLP on the tcgen05 dense
  kernel) -> graph convolutions with [short walk | Q G_s Q^T X | L_e X] messages -> gated readout."""
  L = L.float().contiguous()
  state = embedding(node_ids, model.embedding.weight)
  B, N = state.shape[0], state.shape[1]
  K, S = model.num_eig_vec, model.num_scale_long
  short = list(model.short_diffusion_dist)
  powers = Q = None
  if S > 0:
    adj = (L[:, :, :, 0] != 0).to(torch.float32)                    # ada_lanczos_net.py:310-311
    Le = _gaussian_laplacian_train(state, adj)
    T, Q = _lanczos_train(Le, mask, q1.to(L.device), K)
    plist, cur = [], T
    for p in range(1, max(model.long_diffusion_dist) + 1):          # T^p by repeated products (:262-270)
      if p in model.long_diffusion_dist:
        plist.append(cur)
      if p < max(model.long_diffusion_dist):
        cur = bmm(cur, T)
    powers = torch.cat(plist, dim=2)                                # [B,K,S*K]: index r, s*K + c (:274)
  for t in range(model.num_layer):
    msgs = []
    if short:
      walk = state
      for step in range(1, max(short) + 1):
        walk = operator_messages(L, walk, 0, 1)
        if step in short:
          msgs.append(walk)
    if S > 0:
      if model.spectral_filter_kind == 'MLP':
        h = powers.reshape(B, K * S * K)
        seq = model.spectral_filter[t]
        for i in (0, 2, 4, 6):
          h = dense(h, seq[i].weight, seq[i].bias, i != 6)
        G = h.reshape(B, K, K, S)                                   # index r, c, s (:275)
        G = ((G + G.transpose(1, 2)) * 0.5).permute(0, 3, 1, 2)     # [B,S,K,K]
      else:
        G = torch.stack(plist, dim=1)
      D = state.shape[2]
      U = bmm(Q.transpose(1, 2), state)                             # [B,K,D]
      W = bmm(G.reshape(B * S, K, K), U.unsqueeze(1).expand(B, S, K, D).reshape(B * S, K, D))
      M = bmm(Q.unsqueeze(1).expand(B, S, N, K).reshape(B * S, N, K), W)      # [B*S,N,D]
      msgs.append(M.reshape(B, S, N, D).permute(0, 2, 1, 3).reshape(B, N, S * D))
    msgs.append(operator_messages(L, state))
    lin = model.filter[t]
    state = dense(torch.cat(msgs, dim=2).reshape(B * N, -1), lin.weight, lin.bias, True).reshape(B, N, -1)
    if model.training and model.dropout > 0.0:
      state = torch.nn.functional.dropout(state, model.dropout, True)
  return gated_readout(model, state, mask)


class GraphedStep:
  """One optimisation step -- forward, loss, backward, optimizer update -- of a drop-in module captured
  in ONE CUDA graph and replayed per batch (the loop body of runner/qm8_runner.py:226-259:
  ``optimizer.zero_grad(); _, loss = model(...); loss.backward(); optimizer.step()``).

  At the reference's batch size (64 molecules, config/qm8_lanczos_net.yaml:33) a training step is a few
  hundred small launches and launch-bound in eager mode; the replayed graph removes the host from the
  loop.  Inputs are copied into static device buffers (shapes are fixed at capture: pad every batch to
  the same node count -- padded nodes are masked and have zero operator rows, so the padding does not
  change a real node's value).  The optimizer must support capture (``torch.optim.Adam`` /
  ``AdamW`` get ``capturable=True`` here; plain SGD needs nothing).  The warm-up iterations torch needs
  before capture are rolled back (parameters and optimizer state restored in place), so constructing
  the object does not advance training.  Not for AdaLanczosNet (its start vector is drawn on the
  host each call)."""

  def __init__(self, model, optimizer, args, kwargs=None, warmup=3):
    kwargs = dict(kwargs or {})
    if not hasattr(type(model), '_train_impl') or type(model).__name__ == 'AdaLanczosNet':
      raise TypeError('GraphedStep needs a drop-in module with a host-free training forward')
    if kwargs.get('label') is None:
      raise ValueError('GraphedStep captures the loss: pass label=')
    dev = model._device()
    if dev.type != 'cuda':
      raise RuntimeError('GraphedStep needs the module on a CUDA device')
    model.train()
    self.model, self.optimizer = model, optimizer
    self._args = [self._static(a, dev) for a in args]
    self._kwargs = {k: self._static(v, dev) for k, v in kwargs.items()}
    for group in optimizer.param_groups:
      if 'capturable' in group:
        group['capturable'] = True
    for st in optimizer.state.values():                      # a resumed Adam keeps ``step`` on the host
      if torch.is_tensor(st.get('step')) and st['step'].device != dev:
        st['step'] = st['step'].to(dev)
    params = [p for g in optimizer.param_groups for p in g['params']]
    saved_p = [p.detach().clone() for p in params]
    saved_s = {id(p): {k: v.detach().clone() for k, v in optimizer.state.get(p, {}).items() if torch.is_tensor(v)}
               for p in params}
    side = torch.cuda.Stream(device=dev)
    side.wait_stream(torch.cuda.current_stream(dev))
    with torch.cuda.stream(side):
      for _ in range(max(int(warmup), 1)):
        self._body()
    torch.cuda.current_stream(dev).wait_stream(side)
    self.graph = torch.cuda.CUDAGraph()
    optimizer.zero_grad(set_to_none=True)
    with torch.cuda.graph(self.graph, stream=side):          # same stream as the warm-up: the parameters'
      self.score, self.loss = self._body(zero=False)         # AccumulateGrad nodes were created on it
    with torch.no_grad():                                    # roll the warm-up steps back, in place
      for p, sp in zip(params, saved_p):
        p.copy_(sp)
        for k, v in optimizer.state.get(p, {}).items():
          if torch.is_tensor(v):
            old = saved_s[id(p)].get(k)
            v.copy_(old) if old is not None else v.zero_()
    self.replays = 0

  @staticmethod
  def _static(x, dev):
    return x.detach().to(dev).clone() if torch.is_tensor(x) else x

  def _body(self, zero=True):
    if zero:
      self.optimizer.zero_grad(set_to_none=True)
    score, loss = self.model(*self._args, **self._kwargs)
    loss.backward()
    self.optimizer.step()
    return score, loss

  def __call__(self, *args, **kwargs):
    """Copy this batch into the captured buffers and replay.  Returns (score, loss): static device
    tensors that the next call overwrites."""
    for dst, src in zip(self._args, args):
      if torch.is_tensor(dst):
        if dst.shape != src.shape:
          raise ValueError('GraphedStep was captured for %s, got %s' % (tuple(dst.shape), tuple(src.shape)))
        dst.copy_(src, non_blocking=True)
    for k, dst in self._kwargs.items():
      if torch.is_tensor(dst):
        dst.copy_(kwargs[k], non_blocking=True)
    self.graph.replay()
    self.replays += 1
    return self.score, self.loss
