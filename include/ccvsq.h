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
 *   - calls on one stream keep stream order.  Internally consecutive kernels of this library are chained with
 *     programmatic dependent launch (a kernel's prologue may overlap its predecessor's tail; it waits for the
 *     predecessor before reading any argument buffer) — invisible to the caller, no extra synchronisation needed;
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

#define CCVSQ_VERSION 202 /* major*100 + minor; 2.x: ccvsq_forward_args starts with struct_size; 2.02: ccvsq_peer_* */

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
 *   E_bf16   [ccvsq_codebook_rows(K), D + 16] bf16 out (may be NULL): columns 0..D-1 hold the
 *            RN-rounded code, columns D..D+2 a 3-term BF16 split (hi, mid, lo) of the bias
 *            -0.5*||e_k||^2, the rest zero; padding rows are zero with a bias of -3e38.  The screen
 *            multiplies the 16 extra columns with a constant (1,1,1,0,...) block, so the bias is
 *            added by the tensor core itself.
 *   e_max    [2]     fp32   out (may be NULL): max_k ||e_k||, max_k ||e_k - bf16(e_k)|| (the codebook's own BF16
 *            rounding error, which enters the screening margin)                                  */
int ccvsq_codebook_rows(int K); /* K rounded up so that a sweep with any screen tile width (128 / 96 / 64) fits */
int ccvsq_prepare_codebook(const float* E, int K, int D, float* e_sq, void* E_bf16, float* e_max,
                           void* stream);

/* ---- exact FP32 nearest-code search ------------------------------------------------------
 * Replaces quantize.py:45-50 (distance matrix + argmin) without materialising d[N,K].
 * d = fl(fl(||z||^2 + ||e||^2) - 2*dot) evaluated in FP32, first-occurrence argmin.
 * Works for every D >= 1 (the reference's state quantizer has e_dim = 1, state_model.py:57).
 *   idx [N] int64 out, N = G*S*mult.                                                         */
int ccvsq_search_exact(const float* z, ccvsq_layout lay, const float* E, const float* e_sq, int K,
                       int64_t* idx, void* stream);

/* ---- tensor-core screening + FP32 re-scoring ----------------------------------------------
 * ccvsq_screen: s[n,k] = <bf16(z_n), bf16(e_k)> - 0.5||e_k||^2 on tcgen05 tensor cores (2-CTA MMA,
 * latents resident in tensor memory, FP32 accumulation) with a fused running candidate selection
 * per row.  z is the caller's FP32 tensor in any ccvsq_layout (no packing pass).  A code is a
 * candidate of row n if its score is within
 *     margin_n = margin_tau * 2 * (||z_n - bf16(z_n)|| * e_max[0] * (1 + 2^-8) + ||z_n|| * e_max[1])
 *                + 2^-13 * ||z_n|| * e_max[0]
 * of the row maximum.  With margin_tau = 1 this BOUNDS how far the FP32 winner can trail the BF16 maximum (the error
 * of the difference of two scores whose operands are both rounded to BF16, by Cauchy-Schwarz, plus the FP32
 * accumulation of the tensor core); the two rounding-error norms are measured (per row by the kernel, per codebook
 * by ccvsq_prepare_codebook), so dense operands pay ~1.4 * 2^-8 ||z|| max||e|| and operands on BF16 rounding
 * midpoints up to 4 * 2^-8 ||z|| max||e||.  e_max NULL = (1, 2^-8).
 *   idx         [N] int64 : the only candidate of the row (final), or the best BF16 candidate of a
 *                           queued row (provisional)
 *   queue_count [1] int32, zeroed by the caller: rows queued for ccvsq_rescore, i.e. rows with more
 *                           than one candidate or an incomplete candidate list
 *   queue_rows  [N] int32 : their row numbers
 *   queue_cand  [N, n_cand] int32: their candidates sorted by (score desc, code asc), -1 padded
 *   queue_flags [N] uint8 : bit0 = more than n_cand codes inside the margin (list truncated),
 *                           bit1 = the kernel's internal list overflowed and dropped a code that may
 *                           be inside the margin
 * Requires D % 64 == 0, 64 <= D <= 512, N < 2^31.                                               */
int ccvsq_screen(const float* z, ccvsq_layout lay, const void* E_bf16, const float* e_max, int K,
                 float margin_tau, int n_cand, int64_t* idx, int32_t* queue_count, int32_t* queue_rows,
                 int32_t* queue_cand, uint8_t* queue_flags, void* stream);

/* Diagnostic variant (tests only).  Additionally dumps, for EVERY row, the candidate list
 * cand_idx/cand_score [N, n_cand] (-1 / -inf padded), flags [N], the margin row_margin [N] and, if
 * `scores` is non-null, the full score matrix [N, ccvsq_codebook_rows(K)] (O(N*K) memory).
 * cta_group selects the 2-CTA (2, production) or single-CTA (1) MMA path; idx and the queue pointers
 * may be NULL (together).                                                                       */
int ccvsq_screen_debug(const float* z, ccvsq_layout lay, const void* E_bf16, const float* e_max, int K,
                       float margin_tau, int n_cand, int cta_group, int64_t* idx, int32_t* queue_count,
                       int32_t* queue_rows, int32_t* queue_cand, uint8_t* queue_flags, int32_t* cand_idx,
                       float* cand_score, uint8_t* flags, float* row_margin, float* scores, void* stream);

/* Timeline diagnostic (tools/trace_screen.py): ccvsq_screen plus SM-clock stamps of the pipeline hand-offs.
 * trace [148, 32, 8] int64 (zeroed by the caller): per CTA and row tile (first 32): MMA warp {0: tile start,
 * 1: A operand ready, 2: last MMA issued}, loader warp {3: ready to refill, 4: A buffer free, 5: A stored},
 * epilogue warp {6: first accumulator wait, 7: sweep done}.                                        */
int ccvsq_screen_trace(const float* z, ccvsq_layout lay, const void* E_bf16, const float* e_max, int K,
                       float margin_tau, int n_cand, int64_t* idx, int32_t* queue_count, int32_t* queue_rows,
                       int32_t* queue_cand, uint8_t* queue_flags, int64_t* trace, void* stream);

/* ccvsq_rescore: re-evaluates the candidates of the queued rows in FP32 with the reference's
 * formula and lowest-index tie-break (quantize.py:45-50) and overwrites idx[row].  The queue length
 * is read on the device (no host sync).  Rows with queue_flags != 0 are passed on to the exact
 * fallback:
 *   fallback_ws    int64 [2*fallback_capacity]: row numbers, then packed (distance, code) keys
 *   fallback_count int32 [2], zeroed by the caller: [0] rows queued, [1] scratch counter
 * (both may be NULL: such rows are then resolved among their listed candidates).                */
int ccvsq_rescore(const float* z, ccvsq_layout lay, const float* E, const float* e_sq, int K, int n_cand,
                  const int32_t* queue_count, const int32_t* queue_rows, const int32_t* queue_cand,
                  const uint8_t* queue_flags, int64_t* idx, int64_t* fallback_ws, int32_t* fallback_count,
                  int64_t fallback_capacity, void* stream);

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

/* ---- indices -> prior hand-off -------------------------------------------------
 * The transformer prior starts with  tok_emb(idx) + pos_emb  (mingpt.py:234-236: an nn.Embedding gather
 * of [K, n_embd] rows followed by a broadcast add of the [1, T, n_embd] position table).  Same gather
 * kernel, different table, the add fused into the store:
 *   out[n, :] = fl(table[code[n], :] + pos[n % pos_period, :])      n = b*T + t, pos_period = T
 * Requires D % 4 == 0 and 16-byte aligned pointers; codes outside [0,K) set *err_flag (may be NULL). */
int ccvsq_gather_add(const int64_t* code, const float* table, int K, int D, int64_t N, const float* pos,
                     int64_t pos_period, float* out, int32_t* err_flag, void* stream);

/* ---- Polyak average of the codebook (quantized_video_model.py:951-964, --q_use_ema) ------------
 *   ema[i] = fl(fl(ema[i]*float(decay)) + float(1-decay)*live[i])   in place, one launch (decay is a double like the
 *   Python scalar of the reference, so 1-decay is formed before the cast)
 * (the reference issues par_ema.data.mul_(decay).add_(par.data, alpha=1-decay): two launches per parameter) */
int ccvsq_polyak(float* ema, const float* live, int64_t n, double decay, void* stream);

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

/* Deterministic variant of ccvsq_code_stats (run-to-run bit-identical resid, whatever order the CTAs run in).  The FP32
 * reductions of ccvsq_code_stats / ccvsq_quantize_backward are order-dependent in their last bits; here every term
 * x_n - sub*E[k] is rounded ONCE to a 64-bit fixed-point grid 2^-s (s chosen on the device from max|x|, max|E| and N so
 * that N terms cannot overflow) and accumulated with integer atomics, which are associative.  Three small launches
 * (absmax, accumulate, convert); slower than the fused FP32 path, meant for reproducibility runs.
 *   acc          int64 [K*D] scratch (zeroed by the call)
 *   amax_scratch fp32  [1]   scratch (zeroed by the call)
 *   resid        fp32  [K, D] out (overwritten);  counts int32 [K] (may be NULL), ADDED to            */
int ccvsq_code_stats_fixed(const float* x, ccvsq_layout lay, const float* E, int K, const int64_t* idx,
                           float sub, int64_t* acc, float* amax_scratch, float* resid, int32_t* counts,
                           void* stream);

/* ---- normalize=True (quantize.py:56-57), any mult: assign / backward on the L2-normalised concatenation ----------
 * The quantized vector of a position is u = y/||y||_2 with y the concatenation of its `mult` code rows (norm over all C
 * channels, quantize.py:57).  ccvsq_assign_normalized: z_q = fl(z + fl(u - z)), sq_err += sum (u - z)^2, counts — the
 * same outputs as ccvsq_assign, ready for ccvsq_finalize.  ccvsq_backward_normalized: dz = g_zq + (2 g/M)(z - u) and
 * resid[k,:] = sum_{idx=k} (z - (u.z) u)/||y||  (ZEROED BY THE CALL), so that ccvsq_finalize's
 * dE = -(2 beta g/M) resid is the gradient through the normalisation.  dz or resid may be NULL.                   */
int ccvsq_assign_normalized(const float* z, ccvsq_layout lay, const float* E, int K, const int64_t* idx,
                            float* zq_out, double* sq_err, int32_t* counts, void* stream);
int ccvsq_backward_normalized(const float* z, ccvsq_layout lay, const float* E, int K, const int64_t* idx,
                              const float* g_zq, const float* g_loss, float* dz, float* resid, void* stream);

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
/* Same update from the packed statistics buffer of the multi-GPU exchange, [resid: K*D | counts: K] fp32 (what one
 * all-reduce sums over the ranks; ccvsq_forward_args.resid / .counts_f32 fill it in the forward): no unpacking step. */
int ccvsq_ema_update_packed(float* E, float* n_ema, float* sum_ema, const float* packed, int K, int D,
                            float decay, float eps, float* scratch, void* stream);

/* ---- training collective over NVLink peer memory (EMA extension; replaces the all-reduce of the packed buffer) --------
 * The reference exchanges training state between ranks with NCCL through DDP / apex (tools/engine.py:71-74,127-132).
 * For the EMA statistics [resid: K*D | counts: K] — 1 MB, pure latency for a ring — every rank instead PUSHES its
 * buffer into an inbox slot on every peer (ccvsq_peer_publish, right after the forward; posted NVLink writes + one
 * system-scope flag per peer), and ccvsq_peer_ema_update waits for the W flags, sums the W local slots in rank order
 * (bit-identical on every rank) and applies the EMA update of ccvsq_ema_update in the same pass.  The step counter and
 * the double-buffer parity live in the exchange area, so both calls can be captured in a CUDA graph.
 *   area      ccvsq_peer_exchange_bytes(K, D, world) bytes from ccvsq_peer_alloc (cudaMalloc + cudaIpcGetMemHandle,
 *             zero-filled); the 64-byte handle goes to the other ranks (any transport), which map it with
 *             ccvsq_peer_open.  areas[r] = rank r's area as addressable from THIS process (areas[rank] = own).
 *   Every rank calls publish then update exactly once per step, in that order, on one stream.  A rank that never
 *   publishes makes the others trap after ~60 s (no silent hang, no partial sums).  One process per GPU, one node. */
#define CCVSQ_PEER_MAX_WORLD 16
#define CCVSQ_PEER_HANDLE_BYTES 64
uint64_t ccvsq_peer_exchange_bytes(int K, int D, int world);
int ccvsq_peer_alloc(uint64_t bytes, void** ptr, void* handle64);
int ccvsq_peer_open(const void* handle64, void** ptr);
int ccvsq_peer_close(void* ptr);
int ccvsq_peer_free(void* ptr);
int ccvsq_peer_publish(const float* stats, int K, int D, void* const* areas, int rank, int world, void* stream);
int ccvsq_peer_ema_update(float* E, float* n_ema, float* sum_ema, void* area, int K, int D, int world,
                          float decay, float eps, void* stream);

/* ---- encoder tail: the 1x1 convolution that produces the latents -------------------------
 * Replaces the encoder's last block  ConvLayer(block_out, z_size, 1)  and the optional output normalisation
 * (models/skip_vid_generator/models/skip_autoencoder.py:331,346-349; EqualConv2d :40-58; LeakyReLU(0.1) :98-99):
 *     z[g, o, s] = lrelu_slope( sum_c fl(W[o, c] * scale) * x[g, c, s] + bias[o] ),   then z /= ||z||_2 over o if normalize
 * on tcgen05 tensor cores with FP32-level accuracy: both operands are split into three BF16 terms and the six
 * products above 2^-24 of the result are accumulated in FP32 ("BF16x6").
 *   ccvsq_encoder_tail_prepare: W [C_out, C_in] fp32 -> W_terms [3, C_out, C_in] bf16 (once per weight version)
 *   x [G, C_in, S] fp32 (NCHW, S = h*w), z [G, C_out, S] fp32 out; bias [C_out] (may be NULL)
 * Requires C_in % 64 == 0, C_out % 16 == 0 and (C_out <= 256 or C_out % 256 == 0), 16-byte aligned pointers.        */
int ccvsq_encoder_tail_prepare(const float* W, int C_out, int C_in, float scale, void* W_terms, void* stream);
int ccvsq_encoder_tail(const float* x, int64_t G, int C_in, int S, const void* W_terms, const float* bias,
                       int C_out, float negative_slope, int normalize, float* z, void* stream);

/* ---- whole-op entry points ---------------------------------------------------------------------
 * ccvsq_quantize_forward enqueues the complete forward of the reference module
 * (quantize.py:32-74: flatten, nearest code, gather, loss, straight-through value, perplexity) with
 * ONE call: [prepare_codebook] -> screen -> rescore -> exact fallback (or the exact search) -> assign
 * with the loss / perplexity finalisation folded into its last CTA.  The reference trainer calls the
 * quantizer on 1k-5k latents at a time (scripts/bairhd/train_frame_autoencoder.sh:10,17-20): there the
 * cost is host issue time, which this call keeps to one FFI crossing and one workspace.
 *   header    [CCVSQ_HEADER_INTS + K] int32, caller-owned, ZEROED BY THE CALL; afterwards
 *             header[CCVSQ_HEADER_INTS + k] = per-code usage counts (the rest is scratch)
 *   workspace caller-owned scratch of ccvsq_forward_workspace_bytes(...) bytes (queue arrays and, when
 *             e_sq == NULL, the codebook side data rebuilt on every call)
 *   e_sq / E_bf16 / e_max : cached side data of ccvsq_prepare_codebook (all NULL = build per call);
 *             prepare != 0 rebuilds the cached copy first
 *   idx [N] int64 out; zq (z's layout, may be NULL); loss / perplexity device scalars (may be NULL)
 *   indices_only != 0 stops after the search (QVidModel.encode keeps only info[2],
 *             quantized_video_model.py:798-799)
 *   ev_search_begin / ev_search_end : optional cudaEvent_t recorded around the dominant search kernel
 *             (profiling hook used by bench.py's live roofline timing)                             */
#define CCVSQ_HEADER_INTS 16
#define CCVSQ_SEARCH_AUTO 0
#define CCVSQ_SEARCH_TENSOR 1
#define CCVSQ_SEARCH_EXACT 2

typedef struct ccvsq_forward_args {
  uint32_t struct_size;    /* = sizeof(ccvsq_forward_args) of the header the caller was built against: a binding whose
                              struct layout is stale (a missing trailing field) is rejected with CCVSQ_BAD_SHAPE instead
                              of being read out of bounds */
  uint32_t flags;          /* reserved, must be 0 */
  const float* z;
  ccvsq_layout lay;
  const float* E;
  int32_t K;
  float beta;
  int32_t search_mode;
  int32_t n_cand;
  float margin_tau;
  int32_t exact_fallback;
  int32_t indices_only;
  int32_t prepare;
  float* e_sq;
  void* E_bf16;
  float* e_max;
  int32_t* header;
  void* workspace;
  uint64_t workspace_bytes;
  int64_t* idx;
  float* zq;
  float* loss;
  float* perplexity;
  void* ev_search_begin;
  void* ev_search_end;
  float* resid;            /* optional [K, D] (zeroed by the call): resid[k,:] = sum_{idx=k} (z - E[k]), the per-code
                              statistic of an EMA codebook update, accumulated by the assign pass on its own read of z
                              (no second pass over the latents); NULL = off */
  float* counts_f32;       /* optional [K]: the per-code usage counts written as fp32 by the pass that finalises loss /
                              perplexity (exact below 2^24 per code): the tail of a packed all-reduce buffer; NULL = off */
} ccvsq_forward_args;

uint64_t ccvsq_forward_workspace_bytes(int64_t N, int K, int D, int search_mode, int n_cand,
                                       int with_codebook);
int ccvsq_quantize_forward(const ccvsq_forward_args* args, void* stream);

/* ccvsq_quantize_backward: the autograd backward of quantize.py:55-64 in one pass over z:
 *   dz = g_zq + (2 g_loss / M)(z - E[idx])                       (dz may be NULL)
 *   resid[k,:] = sum_{idx=k} (z - E[k])  (scratch [K, D], zeroed by the call; may be NULL)
 *   dE = -(2 beta g_loss / M) resid                              (may be NULL; may alias resid)
 * The per-code scatter-reduce rides on the same read of z as dz (vector reductions straight from
 * registers) instead of a second pass.                                                          */
int ccvsq_quantize_backward(const float* z, ccvsq_layout lay, const float* E, int K, const int64_t* idx,
                            const float* g_zq, const float* g_loss, float beta, float* dz, float* resid,
                            float* dE, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* CCVSQ_H_ */
