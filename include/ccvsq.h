/*
 * ccvsq.h — C ABI of the B200-native latent vector quantizer (libccvsq.so).
 *
 * This is the drop-in boundary for ONE path of 16lemoing/ccvs: the VectorQuantizer between the
 * frame autoencoder and the transformer prior
 *   reference: models/skip_vid_generator/modules/quantize.py:7-83
 * The reference has no FFI of its own (the boundary is a Python nn.Module); each entry point below
 * cites the reference lines it replaces.  The Python wrapper (ccvs_b200/quantize.py) binds these
 * with ctypes; INTEGRATION.md shows the stub a reference maintainer would add.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer owned by the caller (PyTorch); the library never
 *     allocates, frees or retains device memory;
 *   - every call only ENQUEUES work on `stream` (a cudaStream_t passed as void*); no call
 *     synchronises the device;
 *   - return 0 on success, a negative ccvsq_status otherwise; ccvsq_last_error() gives text;
 *   - no exceptions cross the ABI, no torch types in any signature.
 *
 * Latent layout (ccvsq_layout).  The reference flattens z[B,(T,)C,h,w] to rows by
 *   z.transpose(-3,-1).transpose(-3,-2).contiguous().view(-1, e_dim)      (quantize.py:40-42)
 * We never materialise that copy.  A latent tensor is described by four integers
 *   G    = product of the leading dims (B or B*T)
 *   C    = channels (= e_dim * mult)
 *   S    = h*w (1 when the reference does not transpose, i.e. ndim < 4)
 *   mult = sub-vectors per position (quantize.py:20-21)
 * element (g, c, s) lives at ((g*C + c)*S + s); latent row n = (g*S + s)*mult + m has
 * D = C/mult components j = 0..D-1 taken from channel c = m*D + j.  This is exactly the row
 * order the reference's view(-1, e_dim) produces.
 */
#ifndef CCVSQ_H_
#define CCVSQ_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CCVSQ_VERSION 100 /* major*100 + minor */

typedef enum ccvsq_status {
  CCVSQ_OK = 0,
  CCVSQ_BAD_SHAPE = -1,     /* a size is <= 0, not divisible, or overflows int32 where required */
  CCVSQ_UNSUPPORTED = -2,   /* shape outside what the tensor-core path supports (use the exact path) */
  CCVSQ_MISALIGNED = -3,    /* a pointer violates the stated alignment */
  CCVSQ_CUDA_ERROR = -4,    /* a CUDA runtime/driver call failed; see ccvsq_last_error() */
  CCVSQ_NULL_POINTER = -5
} ccvsq_status;

typedef struct ccvsq_layout {
  int64_t G;    /* leading (batch*time) extent            */
  int32_t C;    /* channels = D * mult                     */
  int32_t S;    /* spatial extent h*w (1 = rows contiguous) */
  int32_t mult; /* sub-vectors per position                */
} ccvsq_layout;

/* Maximum candidates per latent kept by the screening pass. */
#define CCVSQ_MAX_CAND 8

int ccvsq_version(void);
const char* ccvsq_last_error(void); /* thread-local, valid until the next failing call */

/* ---- codebook preparation (once per codebook version) ------------------------------------
 * Replaces the per-call  torch.sum(self.embedding.weight**2, dim=1)  of quantize.py:46 and
 * builds the BF16 shadow used by the tensor-core screening pass.
 *   E        [K, D]  fp32   codebook (embedding.weight, quantize.py:26)
 *   e_sq     [K]     fp32   out: ||e_k||^2
 *   E_bf16   [K_pad, D] bf16 out (may be NULL): RN-rounded copy, rows K..K_pad-1 zero-filled
 *   bias     [K_pad] fp32   out (may be NULL): -0.5*||e_k||^2, -inf for the padding rows
 *   e_max    [1]     fp32   out (may be NULL): max_k ||e_k||
 * K_pad = K rounded up to a multiple of 256.                                                */
int ccvsq_prepare_codebook(const float* E, int K, int D, float* e_sq, void* E_bf16, float* bias,
                           float* e_max, void* stream);

/* ---- exact FP32 nearest-code search ------------------------------------------------------
 * Replaces quantize.py:45-50 (distance matrix + argmin) without materialising d[N,K].
 * d = fl(fl(||z||^2 + ||e||^2) - 2*dot) evaluated in FP32, first-occurrence argmin.
 * Works for every D >= 1 (the reference's state quantizer has e_dim = 1, state_model.py:57).
 *   idx [N] int64 out, N = G*S*mult.                                                         */
int ccvsq_search_exact(const float* z, ccvsq_layout lay, const float* E, const float* e_sq, int K,
                       int64_t* idx, void* stream);

/* ---- tensor-core screening: pack, screen, rescore ----------------------------------------
 * ccvsq_pack_latents: rows of z (any layout) -> row-major BF16 [N_pad, D] for the TMA-fed GEMM,
 * plus the per-row screening margin  margin_scale * ||z_n|| * (*e_max)  (fp32; e_max is the device
 * scalar written by ccvsq_prepare_codebook, NULL = 1).  N_pad = N rounded up to 128; padding rows
 * are zero-filled.  With margin_scale = tau * 2^-8 the margin is tau times the first-order bound
 * on the BF16 rounding error of one score for vectors with evenly spread energy.              */
int ccvsq_pack_latents(const float* z, ccvsq_layout lay, void* z_bf16, float* row_margin,
                       float margin_scale, const float* e_max, void* stream);

/* ccvsq_screen: s[n,k] = <bf16(z_n), bf16(e_k)> - 0.5||e_k||^2 on tcgen05 tensor cores with a
 * fused running candidate selection per row.  The kernel's two epilogue groups each own half of
 * the code tiles (even / odd tiles of 256 codes) and report independently, so every per-row output
 * has two halves h = 0, 1:
 *   cand_idx   [N, 2, n_cand] int32: codes whose score is within row_margin[n] of that half's
 *                                    maximum, sorted by (score desc, code asc), -1 padded
 *   cand_score [N, 2, n_cand] fp32 : their BF16-path scores (-inf padded); slot 0 = half maximum
 *   flags      [N, 2] uint8        : bit0 = more than n_cand codes were inside that half's margin
 *                                    (list truncated), bit1 = the half's internal list overflowed
 *                                    and dropped a code that may be inside the margin
 * ccvsq_rescore merges the halves.
 * Requires D % 64 == 0, 64 <= D <= 512.  z_bf16 is [N_pad, D] (N_pad = N rounded up to 128), E_bf16
 * [K_pad, D] and bias [K_pad] (K_pad = K rounded up to 256), row_margin [N_pad].               */
int ccvsq_screen(const void* z_bf16, const float* row_margin, const void* E_bf16, const float* bias,
                 int64_t N, int K, int D, int n_cand, int32_t* cand_idx, float* cand_score,
                 uint8_t* flags, void* stream);

/* Diagnostic variant: additionally dumps the full score matrix s[N_pad, K_pad] (fp32, row-major,
 * N_pad = N rounded up to 128, K_pad = K rounded up to 256).  Tests only — O(N*K) memory.       */
int ccvsq_screen_dump(const void* z_bf16, const float* row_margin, const void* E_bf16,
                      const float* bias, int64_t N, int K, int D, int n_cand, int32_t* cand_idx,
                      float* cand_score, uint8_t* flags, float* scores, void* stream);

/* Diagnostic variant: CTA 0 records a per-role event timeline into trace (int64 [4][4000], zeroed
 * by the caller; value = clock64 << 8 | event code).  Performance debugging only.               */
int ccvsq_screen_trace(const void* z_bf16, const float* row_margin, const void* E_bf16,
                       const float* bias, int64_t N, int K, int D, int n_cand, int32_t* cand_idx,
                       float* cand_score, uint8_t* flags, long long* trace, void* stream);

/* ccvsq_rescore: merges the two halves of the screen output (a candidate is live if its score is
 * within row_margin[n] of the better half maximum), then re-evaluates the live candidates in FP32
 * with the reference's formula and lowest-index tie-break (quantize.py:45-50); rows with a single
 * live candidate take it directly without touching z.
 * Rows whose candidate set is incomplete w.r.t. the merged threshold (a truncated half whose last
 * slot is still live, or a half that dropped entries while its maximum is live) are queued for the
 * exact fallback:
 *   fallback_ws    int64 [2*fallback_capacity]: row numbers, then packed (distance, code) keys
 *   fallback_count int32 [2], zeroed by the caller: [0] rows queued, [1] scratch counter
 * (both may be NULL to ignore overflow rows).                                                   */
int ccvsq_rescore(const float* z, ccvsq_layout lay, const float* E, const float* e_sq, int K,
                  const int32_t* cand_idx, const float* cand_score, const float* row_margin,
                  int n_cand, const uint8_t* flags, int64_t* idx, int64_t* fallback_ws,
                  int32_t* fallback_count, int64_t fallback_capacity, void* stream);

/* Exact FP32 search restricted to the rows queued by ccvsq_rescore (count read on the device, no
 * host sync).  The codebook is split across CTAs, partial minima meet through 64-bit atomicMin on
 * the packed keys, the last CTA writes idx[row].                                                */
int ccvsq_search_exact_rows(const float* z, ccvsq_layout lay, const float* E, const float* e_sq,
                            int K, int64_t* fallback_ws, int32_t* fallback_count,
                            int64_t fallback_capacity, int64_t* idx, void* stream);

/* ---- assignment: gather + straight-through value + squared error ---------------------------
 * Replaces quantize.py:55 (one-hot GEMM == E[idx]), :60-61 (loss numerator) and :64 (STE value).
 *   zq_out (same layout as z, may be NULL): fl(z + fl(E[idx] - z))   -- the reference's forward
 *          value (bitwise), NOT E[idx]
 *   sq_err [1] fp64 accumulator, ADDED to (caller zeroes): sum (E[idx]-z)^2 (fp32 per-CTA
 *          partials, fp64 only for the cross-CTA atomic so the 1e-5 loss tolerance holds at 2^28
 *          elements)
 *   counts [K] int32 (may be NULL), ADDED to: per-code usage (perplexity, quantize.py:67-68)  */
int ccvsq_assign(const float* z, ccvsq_layout lay, const float* E, int K, const int64_t* idx,
                 float* zq_out, double* sq_err, int32_t* counts, void* stream);

/* ---- decode gather (embed_code, quantize.py:76-83) -----------------------------------------
 * out_lay.S == 1: out[n, :] = E[code[n], :] (row-major, what nn.Embedding returns; the mult>1
 * reshape of quantize.py:78-82 is a pure view of this buffer).
 * out_lay.S  > 1: writes the channel-major [G, C, S] tensor the decoder consumes directly,
 * fusing the caller's NHWC->NCHW copy (quantized_video_model.py:833).
 * Returns CCVSQ_BAD_SHAPE semantics on the device: codes outside [0,K) set *err_flag (int32,
 * may be NULL) instead of reading out of bounds (nn.Embedding raises in that case).           */
int ccvsq_gather(const int64_t* code, const float* E, int K, ccvsq_layout out_lay, float* out,
                 int32_t* err_flag, void* stream);

/* ---- backward: straight-through + commitment gradient ---------------------------------------
 * Replaces the autograd graph of quantize.py:60-64:
 *   dz = g_zq + (2*g_loss/M) * (z - E[idx]),   M = numel(z)
 * g_zq may be NULL (treated as 0); g_loss is a device scalar.                                  */
int ccvsq_backward_dz(const float* z, ccvsq_layout lay, const float* E, int K, const int64_t* idx,
                      const float* g_zq, const float* g_loss, float* dz, void* stream);

/* ---- per-code scatter-reduce ----------------------------------------------------------------
 * resid[k,:] += sum_{n: idx[n]=k} (x_n - sub*E[k]),  counts[k] += |{n: idx[n]=k}|
 * With x = z, sub = 1 this is the sufficient statistic of the codebook gradient
 *   dE[k] = (2*beta*g_loss/M) * (n_k*E_k - sum z) = -(2*beta*g_loss/M) * resid[k]
 * (autograd of quantize.py:55,61); with sub = 0 it is a plain scatter-add of rows (embedding
 * backward / EMA sums).  E may be NULL when sub == 0.  Caller zeroes resid/counts.
 *   resid [K, D] fp32, counts [K] int32 (may be NULL).                                          */
int ccvsq_code_stats(const float* x, ccvsq_layout lay, const float* E, int K, const int64_t* idx,
                     float sub, float* resid, int32_t* counts, void* stream);

/* ---- finalize: codebook gradient, loss, perplexity -------------------------------------------
 *   dE[k,:]   = -(2*beta/M) * g_loss * resid[k,:]           (dE may be NULL)
 *   loss      = (1+beta) * sq_err / M                        (quantize.py:60-61; loss may be NULL)
 *   perplexity= exp(-sum_k p_k log(p_k + 1e-10)), p_k = counts[k]/N   (quantize.py:67-68)
 * M and N are the GLOBAL element / latent counts the statistics were reduced over.             */
int ccvsq_finalize(const float* resid, const int32_t* counts, const double* sq_err,
                   const float* g_loss, int K, int D, double M, double N, float beta, float* dE,
                   float* loss, float* perplexity, void* stream);

/* ---- EMA codebook update (extension; the reference trains the codebook with Adam) --
 *   n_ema   = decay*n_ema + (1-decay)*counts
 *   sum_ema = decay*sum_ema + (1-decay)*(resid + counts*E)
 *   E       = sum_ema / ((n_ema + eps) / (sum(n_ema) + K*eps) * sum(n_ema))
 * scratch: one caller-owned fp32 (receives sum(n_ema)).                                        */
int ccvsq_ema_update(float* E, float* n_ema, float* sum_ema, const float* resid,
                     const int32_t* counts, int K, int D, float decay, float eps, float* scratch,
                     void* stream);

#ifdef __cplusplus
}
#endif
#endif /* CCVSQ_H_ */
