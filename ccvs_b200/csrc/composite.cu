// composite.cu — whole-op entry points: one C call enqueues every kernel of a quantizer forward
// (quantize.py:32-74) or backward (autograd of :55-64).  The real CCVS trainer issues the op on
// 1k-5k latents per call: there the cost is host issue time, so the Python side makes
// ONE ctypes call with one workspace instead of seven calls with a dozen small allocations.
#include "common.cuh"

using namespace ccvsq;

namespace {
inline size_t align_up(size_t x, size_t a = 256) { return (x + a - 1) / a * a; }

struct WsPlan {
  size_t e_sq, e_bf16, q_rows, q_cand, q_flags, fb_ws, total;
  int64_t fb_cap;
};
WsPlan plan_workspace(int64_t N, int K, int D, int n_cand, bool with_codebook, bool tensor) {
  WsPlan p = {};
  size_t off = 0;
  if (with_codebook) {
    p.e_sq = off;   off += align_up((size_t)K * 4);
    p.e_bf16 = off; off += tensor ? align_up((size_t)ccvsq_codebook_rows(K) * (D + SCREEN_EXT) * 2) : 0;
  }
  if (tensor) {
    p.fb_cap = N;   // every row may be flagged (e.g. a collapsed codebook: all codes tie inside the margin)
    p.q_rows = off;  off += align_up((size_t)N * 4);
    p.q_cand = off;  off += align_up((size_t)N * n_cand * 4);
    p.q_flags = off; off += align_up((size_t)N);
    p.fb_ws = off;   off += align_up((size_t)p.fb_cap * 2 * 8);
  }
  p.total = off;
  return p;
}
bool tensor_shape_ok(int K, int D, int64_t N) { return D % 64 == 0 && D >= 64 && D <= 512; }
bool use_tensor_path(int mode, int K, int D, int64_t N) {
  if (mode == CCVSQ_SEARCH_TENSOR) return true;
  return mode == CCVSQ_SEARCH_AUTO && tensor_shape_ok(K, D, N) && N >= 128 && K >= 64;
}
}  // namespace

extern "C" uint64_t ccvsq_forward_workspace_bytes(int64_t N, int K, int D, int search_mode, int n_cand,
                                                  int with_codebook) {
  if (N <= 0 || K <= 0 || D <= 0) return 0;
  return plan_workspace(N, K, D, n_cand, with_codebook != 0, use_tensor_path(search_mode, K, D, N)).total + 256;
}

extern "C" int ccvsq_quantize_forward(const ccvsq_forward_args* a, void* stream) {
  CCVSQ_REQUIRE(a, CCVSQ_NULL_POINTER, "quantize_forward: null args");
  CCVSQ_REQUIRE(a->struct_size == sizeof(ccvsq_forward_args), CCVSQ_BAD_SHAPE,
                "quantize_forward: args.struct_size=%u but this library's ccvsq_forward_args has %zu bytes (ABI %d): the "
                "caller's struct definition is stale", a->struct_size, sizeof(ccvsq_forward_args), CCVSQ_VERSION);
  CCVSQ_REQUIRE(a->flags == 0, CCVSQ_BAD_SHAPE, "quantize_forward: args.flags=%u (reserved, must be 0)", a->flags);
  CCVSQ_REQUIRE(a->z && a->E && a->header && a->idx, CCVSQ_NULL_POINTER, "quantize_forward: z, E, header and idx are required");
  CCVSQ_REQUIRE(a->K > 0, CCVSQ_BAD_SHAPE, "quantize_forward: K=%d", a->K);
  CCVSQ_REQUIRE(a->search_mode >= CCVSQ_SEARCH_AUTO && a->search_mode <= CCVSQ_SEARCH_EXACT, CCVSQ_BAD_SHAPE,
                "quantize_forward: search_mode=%d", a->search_mode);
  Lay L;
  if (int rc = make_lay(a->lay, &L)) return rc;
  const int K = a->K, D = L.D;
  cudaStream_t st = (cudaStream_t)stream;
  const bool tensor = use_tensor_path(a->search_mode, K, D, L.N);
  const bool own_cb = a->e_sq == nullptr;
  CCVSQ_REQUIRE(own_cb || !tensor || (a->E_bf16 && a->e_max), CCVSQ_NULL_POINTER,
                "quantize_forward: cached codebook side data lacks the BF16 shadow / e_max");
  const WsPlan p = plan_workspace(L.N, K, D, a->n_cand, own_cb, tensor);
  uint8_t* ws = (uint8_t*)(((uintptr_t)a->workspace + 255) & ~(uintptr_t)255);
  CCVSQ_REQUIRE(p.total == 0 || (a->workspace && ws + p.total <= (uint8_t*)a->workspace + a->workspace_bytes),
                CCVSQ_BAD_SHAPE, "quantize_forward: workspace of %llu bytes is too small (need %llu)",
                (unsigned long long)a->workspace_bytes, (unsigned long long)(p.total + 256));

  // header: [0] queue count | [1,2] fallback counters | [3] ticket | [4,5] fp64 sum of squared errors |
  //         [6] max ||e||, [7] max ||e - bf16(e)|| (when the side data is built here) | [16, 16+K) per-code counts
  int32_t* hdr = a->header;
  CCVSQ_CUDA(cudaMemsetAsync(hdr, 0, (size_t)(CCVSQ_HEADER_INTS + K) * sizeof(int32_t), st));
  // (all memsets come first: the kernels below form one programmatic-dependent-launch chain)
  if (a->resid && !a->indices_only) CCVSQ_CUDA(cudaMemsetAsync(a->resid, 0, (size_t)K * D * sizeof(float), st));
  int32_t* q_count = hdr + 0;
  int32_t* fb_count = hdr + 1;
  int32_t* ticket = hdr + 3;
  double* sq_err = reinterpret_cast<double*>(hdr + 4);
  int32_t* counts = hdr + CCVSQ_HEADER_INTS;

  float* e_sq = a->e_sq;
  void* e_bf16 = a->E_bf16;
  float* e_max = a->e_max;
  if (own_cb) {
    e_sq = (float*)(ws + p.e_sq);
    e_bf16 = tensor ? (void*)(ws + p.e_bf16) : nullptr;
    e_max = tensor ? reinterpret_cast<float*>(hdr + 6) : nullptr;
  }
  if (own_cb || a->prepare) {
    if (e_max && !own_cb) CCVSQ_CUDA(cudaMemsetAsync(e_max, 0, 2 * sizeof(float), st));
    if (int rc = prepare_codebook_launch(a->E, K, D, e_sq, e_bf16, e_max, st)) return rc;
  }

  if (a->ev_search_begin) CCVSQ_CUDA(cudaEventRecord((cudaEvent_t)a->ev_search_begin, st));
  if (tensor) {
    int32_t* q_rows = (int32_t*)(ws + p.q_rows);
    int32_t* q_cand = (int32_t*)(ws + p.q_cand);
    uint8_t* q_flags = ws + p.q_flags;
    int64_t* fb_ws = (int64_t*)(ws + p.fb_ws);
    if (int rc = screen_launch_stable_z(a->z, a->lay, e_bf16, e_max, K, a->margin_tau, a->n_cand, a->idx, q_count, q_rows,
                                        q_cand, q_flags, stream))
      return rc;
    if (a->ev_search_end) CCVSQ_CUDA(cudaEventRecord((cudaEvent_t)a->ev_search_end, st));
    const bool fb = a->exact_fallback != 0;
    if (int rc = ccvsq_rescore(a->z, a->lay, a->E, e_sq, K, a->n_cand, q_count, q_rows, q_cand, q_flags, a->idx,
                               fb ? fb_ws : nullptr, fb ? fb_count : nullptr, p.fb_cap, stream))
      return rc;
    if (fb)
      if (int rc = ccvsq_search_exact_rows(a->z, a->lay, a->E, e_sq, K, fb_ws, fb_count, p.fb_cap, a->idx, stream))
        return rc;
  } else {
    if (int rc = ccvsq_search_exact(a->z, a->lay, a->E, e_sq, K, a->idx, stream)) return rc;
    if (a->ev_search_end) CCVSQ_CUDA(cudaEventRecord((cudaEvent_t)a->ev_search_end, st));
  }

  if (a->indices_only) return CCVSQ_OK;
  StreamArgs s = {};
  s.x = a->z; s.E = a->E; s.idx = a->idx; s.out = a->zq; s.sq_err = sq_err; s.counts = counts; s.K = K;
  s.x_stable = true;      // z was complete before prepare_codebook (an ordinary launch) started
  s.fin.ticket = ticket;
  s.fin.loss = a->loss;
  s.fin.perplexity = a->perplexity;
  s.fin.counts_f32 = a->counts_f32;
  s.fin.M = (double)L.P * L.C;
  s.fin.N = (double)L.N;
  s.fin.beta = a->beta;
  if (a->resid) {   // per-code residual sums on the same pass (EMA codebook update; zeroed at the top of the call)
    s.resid = a->resid;
    s.sub = 1.f;
  }
  return stream_launch(MODE_ASSIGN, s, L, st);
}

extern "C" int ccvsq_quantize_backward(const float* z, ccvsq_layout lay, const float* E, int K, const int64_t* idx,
                                       const float* g_zq, const float* g_loss, float beta, float* dz, float* resid,
                                       float* dE, void* stream) {
  CCVSQ_REQUIRE(z && E && idx && g_loss, CCVSQ_NULL_POINTER, "quantize_backward: null pointer");
  CCVSQ_REQUIRE(K > 0, CCVSQ_BAD_SHAPE, "quantize_backward: K=%d", K);
  CCVSQ_REQUIRE(!dE || resid, CCVSQ_NULL_POINTER, "quantize_backward: dE needs the resid scratch");
  if (!dz && !resid) return CCVSQ_OK;
  Lay L;
  if (int rc = make_lay(lay, &L)) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  if (resid) CCVSQ_CUDA(cudaMemsetAsync(resid, 0, (size_t)K * L.D * sizeof(float), st));
  StreamArgs s = {};
  s.x = z; s.g = g_zq; s.E = E; s.idx = idx; s.out = dz; s.resid = resid; s.g_loss = g_loss; s.K = K;
  s.coef_scale = (float)(2.0 / ((double)L.P * L.C));
  s.sub = 1.f;
  if (int rc = stream_launch(dz ? MODE_BACKWARD : MODE_STATS, s, L, st)) return rc;
  if (dE)   // dE = -(2 beta g_loss / M) resid; dE may alias resid (elementwise)
    return ccvsq_finalize(resid, nullptr, nullptr, g_loss, K, L.D, (double)L.P * L.C, (double)L.N, beta, dE, nullptr,
                          nullptr, stream);
  return CCVSQ_OK;
}
