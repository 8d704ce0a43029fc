// torch custom-op registration over the C ABI (include/tcct_b200.h): `torch.ops.tcct_b200.*`.
//
// The reference is pure PyTorch: the interface a maintainer binds the kernels through is the PyTorch operator surface
// (SURVEY 8b).  The kernels themselves live in libtcct_b200.so (no torch dependency); this shim only unwraps at::Tensor
// arguments into pointers + sizes, takes the current CUDA stream, allocates outputs through the caching allocator and turns
// a non-zero status into a RuntimeError carrying tcct_last_error().  Schema strings below are the operator ABI.
#include <ATen/ATen.h>
#include <c10/cuda/CUDAStream.h>
#include <torch/library.h>

#include "tcct_b200.h"

namespace {

void* stream_of(const at::Tensor& t) { return (void*)c10::cuda::getCurrentCUDAStream(t.device().index()).stream(); }

void check(int rc, const char* what) {
  TORCH_CHECK(rc == 0, "tcct_b200::", what, " failed (", rc, "): ", tcct_last_error());
}
void cuda_f32(const at::Tensor& t, const char* name) {
  TORCH_CHECK(t.is_cuda() && t.is_contiguous() && t.scalar_type() == at::kFloat, "tcct_b200: `", name,
              "` must be a contiguous float32 CUDA tensor (there is no CPU path)");
}
const float* fptr(const c10::optional<at::Tensor>& t) { return t.has_value() && t->defined() ? t->data_ptr<float>() : nullptr; }

// 32 -> 32 channel 3x3 / 1xk / kx1 convolution, NHWC, weights packed by tcct_pack_weights (fmt 2).  nets/tcct.py:803-828
at::Tensor conv2d_tma(const at::Tensor& x, const at::Tensor& w_packed, const c10::optional<at::Tensor>& bias, int64_t KH, int64_t KW,
                      const c10::optional<at::Tensor>& stats, int64_t stats_act) {
  cuda_f32(x, "x"); cuda_f32(w_packed, "w_packed");
  TORCH_CHECK(x.dim() == 4 && x.size(3) == 32, "conv2d_tma: x must be NHWC [B,H,W,32]");
  at::Tensor y = at::empty_like(x);
  double* st = stats.has_value() && stats->defined() ? stats->data_ptr<double>() : nullptr;
  check(tcct_conv2d_tma(x.data_ptr<float>(), w_packed.data_ptr<float>(), fptr(bias), y.data_ptr<float>(), (int)x.size(0), (int)x.size(1),
                        (int)x.size(2), (int)KH, (int)KW, st, (int)stats_act, stream_of(x)), "conv2d_tma");
  return y;
}

// weight / bias gradient of the same convolutions, accumulated into dw [Cout,32,KH,KW] / dbias [Cout]
void conv2d_wgrad_tma(const at::Tensor& x, const at::Tensor& dy, at::Tensor dw, c10::optional<at::Tensor> dbias) {
  cuda_f32(x, "x"); cuda_f32(dy, "dy"); cuda_f32(dw, "dw");
  const int B = (int)x.size(0), H = (int)x.size(1), W = (int)x.size(2), KH = (int)dw.size(2), KW = (int)dw.size(3);
  at::Tensor ws = at::empty({tcct_wgrad_tma_ws_floats(B, H, W, KH, KW)}, x.options());
  float* db = dbias.has_value() && dbias->defined() ? dbias->data_ptr<float>() : nullptr;
  check(tcct_wgrad_tma(x.data_ptr<float>(), dy.data_ptr<float>(), dw.data_ptr<float>(), db, B, H, W, KH, KW, (int)dw.size(0),
                       ws.data_ptr<float>(), nullptr, stream_of(x)), "conv2d_wgrad_tma");
}

// deep-supervision Dice over [z0 | native-resolution auxiliary logits]: kite/loopback.py:62-73, kite/losses/loss.py:83-99
std::tuple<at::Tensor, at::Tensor> dice_multi_fwd(const at::Tensor& z0, const at::Tensor& z1, const at::Tensor& z2, const at::Tensor& z3,
                                                  const at::Tensor& lab, double w_aux) {
  cuda_f32(z0, "z0"); cuda_f32(z1, "z1"); cuda_f32(z2, "z2"); cuda_f32(z3, "z3");
  TORCH_CHECK(lab.is_cuda() && lab.is_contiguous() && lab.scalar_type() == at::kByte, "dice_multi_fwd: lab must be a uint8 CUDA label map");
  const int B = (int)z0.size(0), C = (int)z0.size(1), H = (int)z0.size(2), W = (int)z0.size(3);
  const int hs[3] = {(int)z1.size(2), (int)z2.size(2), (int)z3.size(2)}, ws[3] = {(int)z1.size(3), (int)z2.size(3), (int)z3.size(3)};
  const float wt[4] = {1.f, (float)w_aux, (float)w_aux, (float)w_aux};
  at::Tensor sums = at::zeros({tcct_dice_multi_sums_doubles(C)}, z0.options().dtype(at::kDouble));
  at::Tensor loss = at::empty({5}, z0.options()), coef = at::empty({8 * C}, z0.options());
  check(tcct_dice_multi_fwd(z0.data_ptr<float>(), z1.data_ptr<float>(), z2.data_ptr<float>(), z3.data_ptr<float>(), hs, ws,
                            lab.data_ptr<uint8_t>(), B, C, H, W, wt, sums.data_ptr<double>(), loss.data_ptr<float>(), coef.data_ptr<float>(),
                            stream_of(z0)), "dice_multi_fwd");
  return {loss, coef};
}

std::tuple<at::Tensor, at::Tensor, at::Tensor, at::Tensor> dice_multi_bwd(const at::Tensor& z0, const at::Tensor& z1, const at::Tensor& z2,
                                                                          const at::Tensor& z3, const at::Tensor& lab, double w_aux,
                                                                          const at::Tensor& coef, const at::Tensor& gscale) {
  cuda_f32(z0, "z0"); cuda_f32(coef, "coef"); cuda_f32(gscale, "gscale");
  const int B = (int)z0.size(0), C = (int)z0.size(1), H = (int)z0.size(2), W = (int)z0.size(3);
  const int hs[3] = {(int)z1.size(2), (int)z2.size(2), (int)z3.size(2)}, ws[3] = {(int)z1.size(3), (int)z2.size(3), (int)z3.size(3)};
  const float wt[4] = {1.f, (float)w_aux, (float)w_aux, (float)w_aux};
  at::Tensor d0 = at::empty_like(z0), d1 = at::zeros_like(z1), d2 = at::zeros_like(z2), d3 = at::zeros_like(z3);
  check(tcct_dice_multi_bwd(z0.data_ptr<float>(), z1.data_ptr<float>(), z2.data_ptr<float>(), z3.data_ptr<float>(), hs, ws,
                            lab.data_ptr<uint8_t>(), B, C, H, W, wt, coef.data_ptr<float>(), gscale.data_ptr<float>(), d0.data_ptr<float>(),
                            d1.data_ptr<float>(), d2.data_ptr<float>(), d3.data_ptr<float>(), stream_of(z0)), "dice_multi_bwd");
  return {d0, d1, d2, d3};
}

// KiteSeg.predict: argmax label map of NCHW logits (kite/loop_seg.py:21-33)
at::Tensor argmax_labels(const at::Tensor& logits) {
  cuda_f32(logits, "logits");
  at::Tensor lab = at::empty({logits.size(0), logits.size(2), logits.size(3)}, logits.options().dtype(at::kByte));
  check(tcct_argmax_nchw(logits.data_ptr<float>(), lab.data_ptr<uint8_t>(), (int)logits.size(0), (int)logits.size(1),
                         (int)(logits.size(2) * logits.size(3)), stream_of(logits)), "argmax_labels");
  return lab;
}

// nets/reg.py:27-35
at::Tensor soft_argmax(const at::Tensor& logits, double beta) {
  cuda_f32(logits, "logits");
  at::Tensor out = at::empty({logits.size(0), 1, logits.size(2), logits.size(3)}, logits.options());
  check(tcct_soft_argmax(logits.data_ptr<float>(), out.data_ptr<float>(), (int)logits.size(0), (int)logits.size(1),
                         (int)(logits.size(2) * logits.size(3)), (float)beta, stream_of(logits)), "soft_argmax");
  return out;
}

// soft-argmax boundary extraction
at::Tensor boundary_positions(const at::Tensor& logits, double beta) {
  cuda_f32(logits, "logits");
  at::Tensor out = at::empty({logits.size(0), logits.size(1) - 1, logits.size(3)}, logits.options());
  check(tcct_boundary_positions(logits.data_ptr<float>(), out.data_ptr<float>(), (int)logits.size(0), (int)logits.size(1), (int)logits.size(2),
                                (int)logits.size(3), (float)beta, stream_of(logits)), "boundary_positions");
  return out;
}

// validation scores: kite/losses/miou.py:28-44,69-91
at::Tensor score_sums(const at::Tensor& pr, const at::Tensor& gt) {
  cuda_f32(pr, "pr");
  TORCH_CHECK(gt.is_cuda() && gt.is_contiguous() && gt.sizes() == pr.sizes() && (gt.scalar_type() == at::kFloat || gt.scalar_type() == at::kLong),
              "score_sums: gt must be a float32 or int64 CUDA tensor of pr's shape");
  at::Tensor out = at::zeros({pr.size(0), pr.size(1), 3}, pr.options().dtype(at::kDouble));
  check(tcct_score_sums(pr.data_ptr<float>(), gt.data_ptr(), gt.scalar_type() == at::kLong, (int)pr.size(0), (int)pr.size(1),
                        (int)(pr.size(2) * pr.size(3)), out.data_ptr<double>(), stream_of(pr)), "score_sums");
  return out;
}

int64_t route_count(int64_t id) { return tcct_route_count((int)id); }

// MHCABlock token mixer (nets/tcct.py:457-469): LayerNorm1 -> MetaPool -> residual (DropPath scale) -> LayerNorm2.  SURVEY 8(b) ln_metapool
std::tuple<at::Tensor, at::Tensor, at::Tensor> ln_metapool_fwd(const at::Tensor& t, const at::Tensor& g1, const at::Tensor& b1, const at::Tensor& g2,
                                                               const at::Tensor& b2, const c10::optional<at::Tensor>& scale, double eps) {
  cuda_f32(t, "t"); cuda_f32(g1, "g1"); cuda_f32(b1, "b1"); cuda_f32(g2, "g2"); cuda_f32(b2, "b2");
  TORCH_CHECK(t.dim() == 3, "ln_metapool_fwd: t must be [B,N,C]");
  const int B = (int)t.size(0), N = (int)t.size(1), C = (int)t.size(2);
  at::Tensor t2 = at::empty_like(t), cur2 = at::empty_like(t), stats = at::empty({(int64_t)B * N * 4}, t.options());
  check(tcct_ln_metapool_fwd(t.data_ptr<float>(), g1.data_ptr<float>(), b1.data_ptr<float>(), g2.data_ptr<float>(), b2.data_ptr<float>(), fptr(scale),
                             t2.data_ptr<float>(), cur2.data_ptr<float>(), stats.data_ptr<float>(), B, N, C, (float)eps, stream_of(t)), "ln_metapool_fwd");
  return {t2, cur2, stats};
}

// returns dt; dg1, db1, dg2, db2 [C] are accumulated in place
at::Tensor ln_metapool_bwd(const at::Tensor& t, const at::Tensor& t2, const at::Tensor& stats, const at::Tensor& g1, const at::Tensor& g2,
                           const c10::optional<at::Tensor>& scale, const c10::optional<at::Tensor>& dt2, const c10::optional<at::Tensor>& dcur2,
                           at::Tensor dg1, at::Tensor db1, at::Tensor dg2, at::Tensor db2) {
  cuda_f32(t, "t"); cuda_f32(t2, "t2"); cuda_f32(stats, "stats"); cuda_f32(dg1, "dg1"); cuda_f32(db1, "db1"); cuda_f32(dg2, "dg2"); cuda_f32(db2, "db2");
  const int B = (int)t.size(0), N = (int)t.size(1), C = (int)t.size(2);
  at::Tensor dt = at::empty_like(t);
  check(tcct_ln_metapool_bwd(t.data_ptr<float>(), t2.data_ptr<float>(), stats.data_ptr<float>(), g1.data_ptr<float>(), g2.data_ptr<float>(), fptr(scale),
                             fptr(dt2), fptr(dcur2), dt.data_ptr<float>(), dg1.data_ptr<float>(), db1.data_ptr<float>(), dg2.data_ptr<float>(),
                             db2.data_ptr<float>(), B, N, C, stream_of(t)), "ln_metapool_bwd");
  return dt;
}

// clip_grad_norm_(max_norm) + AdamW over flat buffers (kite/loop_seg.py:128-130, kite/loopback.py:126-128).  SURVEY 8(b) clip_adamw_step
// state: float[4] on the device = [step, lr, last grad norm, -]; p, m, v are updated in place; returns nothing.
void clip_adamw_step(at::Tensor p, const at::Tensor& g, at::Tensor m, at::Tensor v, at::Tensor state, double max_norm, double beta1, double beta2,
                     double eps, double weight_decay, double grad_scale) {
  cuda_f32(p, "p"); cuda_f32(g, "g"); cuda_f32(m, "m"); cuda_f32(v, "v"); cuda_f32(state, "state");
  TORCH_CHECK(g.numel() == p.numel() && m.numel() == p.numel() && v.numel() == p.numel() && state.numel() >= 4, "clip_adamw_step: size mismatch");
  at::Tensor sq = at::zeros({2}, p.options().dtype(at::kDouble));
  check(tcct_sqnorm(g.data_ptr<float>(), (long long)g.numel(), sq.data_ptr<double>(), stream_of(p)), "sqnorm");
  check(tcct_adamw_step(p.data_ptr<float>(), g.data_ptr<float>(), m.data_ptr<float>(), v.data_ptr<float>(), (long long)p.numel(), sq.data_ptr<double>(),
                        state.data_ptr<float>(), (float)max_norm, (float)beta1, (float)beta2, (float)eps, (float)weight_decay, (float)grad_scale,
                        stream_of(p)), "adamw_step");
}

// GateFusion (nets/tcct.py:916-932), NHWC operands, alpha = the small [B,C,hs,ws] random field or None (eval: 0.5)
at::Tensor gate_fuse_fwd(const at::Tensor& x1, const at::Tensor& x2, const c10::optional<at::Tensor>& alpha) {
  cuda_f32(x1, "x1"); cuda_f32(x2, "x2");
  TORCH_CHECK(x1.dim() == 4 && x1.sizes() == x2.sizes(), "gate_fuse_fwd: x1, x2 must be NHWC tensors of one shape");
  const bool has = alpha.has_value() && alpha->defined();
  if (has) cuda_f32(*alpha, "alpha");
  at::Tensor out = at::empty_like(x1);
  check(tcct_gate_fuse_fwd(x1.data_ptr<float>(), x2.data_ptr<float>(), fptr(alpha), out.data_ptr<float>(), (int)x1.size(0), (int)x1.size(1), (int)x1.size(2),
                           (int)x1.size(3), has ? (int)alpha->size(2) : 0, has ? (int)alpha->size(3) : 0, stream_of(x1)), "gate_fuse_fwd");
  return out;
}
std::tuple<at::Tensor, at::Tensor> gate_fuse_bwd(const at::Tensor& dy, const c10::optional<at::Tensor>& alpha) {
  cuda_f32(dy, "dy");
  const bool has = alpha.has_value() && alpha->defined();
  at::Tensor d1 = at::empty_like(dy), d2 = at::empty_like(dy);
  check(tcct_gate_fuse_bwd(dy.data_ptr<float>(), fptr(alpha), d1.data_ptr<float>(), d2.data_ptr<float>(), (int)dy.size(0), (int)dy.size(1), (int)dy.size(2),
                           (int)dy.size(3), has ? (int)alpha->size(2) : 0, has ? (int)alpha->size(3) : 0, stream_of(dy)), "gate_fuse_bwd");
  return {d1, d2};
}

// readPair + make_tran + tensor conversion (data/octnpy.py:117-129, data/octgen.py:9-19,117-126): uint8 frames [B,Hs,Ws,3], gray-level
// labels [B,Hs,Ws], params = B records of tcct_aug_params_size() bytes (uint8 [B, size]) -> (float [B,3,H,W], uint8 [B,H,W])
std::tuple<at::Tensor, at::Tensor> prep_augment(const at::Tensor& img, const at::Tensor& lab, const at::Tensor& params, int64_t row0, int64_t rows,
                                                int64_t Hp, int64_t Wp, int64_t H, int64_t W, int64_t divide) {
  TORCH_CHECK(img.is_cuda() && img.is_contiguous() && img.scalar_type() == at::kByte && img.dim() == 4 && img.size(3) == 3, "prep_augment: img must be uint8 CUDA [B,Hs,Ws,3]");
  TORCH_CHECK(lab.is_cuda() && lab.is_contiguous() && lab.scalar_type() == at::kByte && lab.dim() == 3, "prep_augment: lab must be uint8 CUDA [B,Hs,Ws]");
  TORCH_CHECK(params.is_cuda() && params.is_contiguous() && params.scalar_type() == at::kByte && params.numel() == img.size(0) * tcct_aug_params_size(),
              "prep_augment: params must hold one ", tcct_aug_params_size(), "-byte record per frame");
  at::Tensor out = at::empty({img.size(0), 3, H, W}, img.options().dtype(at::kFloat)), ol = at::empty({img.size(0), H, W}, img.options());
  check(tcct_prep_augment(img.data_ptr<uint8_t>(), lab.data_ptr<uint8_t>(), params.data_ptr(), (int)img.size(0), (int)img.size(1), (int)img.size(2), (int)row0,
                          (int)rows, (int)Hp, (int)Wp, (int)H, (int)W, (int)divide, out.data_ptr<float>(), ol.data_ptr<uint8_t>(), stream_of(img)), "prep_augment");
  return {out, ol};
}

}  // namespace

TORCH_LIBRARY(tcct_b200, m) {
  m.def("conv2d_tma(Tensor x, Tensor w_packed, Tensor? bias, int KH, int KW, Tensor? stats, int stats_act) -> Tensor");
  m.def("conv2d_wgrad_tma(Tensor x, Tensor dy, Tensor(a!) dw, Tensor(b!)? dbias) -> ()");
  m.def("dice_multi_fwd(Tensor z0, Tensor z1, Tensor z2, Tensor z3, Tensor lab, float w_aux) -> (Tensor, Tensor)");
  m.def("dice_multi_bwd(Tensor z0, Tensor z1, Tensor z2, Tensor z3, Tensor lab, float w_aux, Tensor coef, Tensor gscale) -> (Tensor, Tensor, Tensor, Tensor)");
  m.def("argmax_labels(Tensor logits) -> Tensor");
  m.def("soft_argmax(Tensor logits, float beta) -> Tensor");
  m.def("boundary_positions(Tensor logits, float beta) -> Tensor");
  m.def("score_sums(Tensor pr, Tensor gt) -> Tensor");
  m.def("route_count(int id) -> int", &route_count);
  m.def("ln_metapool_fwd(Tensor t, Tensor g1, Tensor b1, Tensor g2, Tensor b2, Tensor? scale, float eps) -> (Tensor, Tensor, Tensor)");
  m.def("ln_metapool_bwd(Tensor t, Tensor t2, Tensor stats, Tensor g1, Tensor g2, Tensor? scale, Tensor? dt2, Tensor? dcur2, Tensor(a!) dg1, "
        "Tensor(b!) db1, Tensor(c!) dg2, Tensor(d!) db2) -> Tensor");
  m.def("clip_adamw_step(Tensor(a!) p, Tensor g, Tensor(b!) m, Tensor(c!) v, Tensor(d!) state, float max_norm, float beta1, float beta2, float eps, "
        "float weight_decay, float grad_scale) -> ()");
  m.def("gate_fuse_fwd(Tensor x1, Tensor x2, Tensor? alpha) -> Tensor");
  m.def("gate_fuse_bwd(Tensor dy, Tensor? alpha) -> (Tensor, Tensor)");
  m.def("prep_augment(Tensor img, Tensor lab, Tensor params, int row0, int rows, int Hp, int Wp, int H, int W, int divide) -> (Tensor, Tensor)");
}

TORCH_LIBRARY_IMPL(tcct_b200, CUDA, m) {
  m.impl("conv2d_tma", &conv2d_tma);
  m.impl("conv2d_wgrad_tma", &conv2d_wgrad_tma);
  m.impl("dice_multi_fwd", &dice_multi_fwd);
  m.impl("dice_multi_bwd", &dice_multi_bwd);
  m.impl("argmax_labels", &argmax_labels);
  m.impl("soft_argmax", &soft_argmax);
  m.impl("boundary_positions", &boundary_positions);
  m.impl("score_sums", &score_sums);
  m.impl("ln_metapool_fwd", &ln_metapool_fwd);
  m.impl("ln_metapool_bwd", &ln_metapool_bwd);
  m.impl("clip_adamw_step", &clip_adamw_step);
  m.impl("gate_fuse_fwd", &gate_fuse_fwd);
  m.impl("gate_fuse_bwd", &gate_fuse_bwd);
  m.impl("prep_augment", &prep_augment);
}
