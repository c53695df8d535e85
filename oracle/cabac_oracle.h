/* TEST INFRASTRUCTURE ONLY -- never linked, imported or called by the product path.
 *
 * CPU restatement (plain C11) of the reference's CABAC hot path, used as the
 * parity oracle by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline
 * leg.  Every function cites the reference file:line it restates (paths relative
 * to /root/reference).  PINNED: checked against the known-answer vectors K0-K8
 *  in tests/golden/kat.json and, when oracle/_ref is built,
 * against the unmodified reference engine on millions of random bins
 * (tests/test_oracle_vs_ref.py).  The MATLAB-level pieces (binarizer, finish
 * detector, context selection) have no golden vectors in the reference; they
 * are pinned by round trips through the reference engine and by the code tables
 * in SURVEY.md 8(a) row a20 (see DESIGN.md "Oracle").
 */
#ifndef ISSCABAC_ORACLE_H
#define ISSCABAC_ORACLE_H
#include <stdint.h>
#include <stddef.h>
#ifdef __cplusplus
extern "C" {
#endif

/* ---- op format (same as include/isscabac.h) ------------------------------ */
#define ORC_OP8_TRM 125u
#define ORC_OP8_EP 126u
#define ORC_OP16_TRM 0x7FFDu
#define ORC_OP16_EP 0x7FFEu

/* ---- binarization methods (cabacBinarizer.m:12-27) ----------------------- */
enum { ORC_BIN_TU = 0, ORC_BIN_EG0 = 1, ORC_BIN_EG1 = 2, ORC_BIN_EG2 = 3, ORC_BIN_FL32 = 4,
       ORC_BIN_TR0 = 5, ORC_BIN_TR1 = 6, ORC_BIN_TR2 = 7 };

/* ---- context-selection profiles ------------------------------------------ */
enum { ORC_PROFILE_DEMO = 0,      /* cabacDemo.m:113-121, 3 contexts                   */
       ORC_PROFILE_ISS = 1,       /* cabacContextSelection.m:24-67, 7*Nlbp+2 contexts  */
       ORC_PROFILE_FLAT = 2,      /* neighbour-free restriction, 2*Nlbp+2 contexts     */
       ORC_PROFILE_FLAT_EPSUF = 3 /* as FLAT but suffix bins are bypass coded          */ };
/* ISS cmTypes mask (cabacContextSelection.m:40-46,58-62) */
enum { ORC_CM_COND0 = 1, ORC_CM_COND1 = 2, ORC_CM_CONDBINLFT = 4, ORC_CM_CONDS0 = 8, ORC_CM_CONDS1 = 16 };

/* ---- context model (ContextModel.h:78-85, ContextModel.cpp:70-74,107-123) - */
uint8_t orc_ctx_make(unsigned mps, unsigned state);
uint8_t orc_ctx_next_mps(uint8_t s);
uint8_t orc_ctx_next_lps(uint8_t s);
unsigned orc_lps_range(unsigned state, unsigned quartile); /* Encoder.cpp:414-480 */
unsigned orc_renorm_shift(unsigned lps);                   /* Encoder.cpp:482-492 */

/* ---- encoder (CABAC_ArithmeticEncoder.cpp:54-412 + BitstreamFile.cpp:95-151) */
typedef struct {
  uint32_t low, range;
  int32_t bits_left;
  uint32_t buffered_byte;
  int32_t num_buffered;
  uint64_t bins_coded;
  /* bit sink */
  uint8_t* out;
  uint64_t cap, n_out; /* n_out counts every byte, even beyond cap */
  uint32_t held, n_held;
  uint64_t bits_written;
} orc_encoder;
void orc_enc_attach(orc_encoder* e, uint8_t* out, uint64_t cap);
void orc_enc_start(orc_encoder* e);
void orc_enc_bin(orc_encoder* e, unsigned bin, uint8_t* ctx);
void orc_enc_ep(orc_encoder* e, unsigned bin);
void orc_enc_bins_ep(orc_encoder* e, uint32_t bins, int n);
void orc_enc_trm(orc_encoder* e, unsigned bin);
void orc_enc_finish(orc_encoder* e);
uint64_t orc_enc_num_bits(const orc_encoder* e); /* CABAC_BitstreamFile.h:70 */

/* ---- decoder (CABAC_ArithmeticDecoder.cpp:54-472 + BitstreamFile.cpp:153-158) */
typedef struct {
  uint32_t range, value;
  int32_t bits_needed;
  const uint8_t* in;
  uint64_t len, pos; /* pos may exceed len: reads past the end give 0xFF */
  uint32_t last_byte;
} orc_decoder;
void orc_dec_start(orc_decoder* d, const uint8_t* in, uint64_t len);
unsigned orc_dec_bin(orc_decoder* d, uint8_t* ctx);
unsigned orc_dec_ep(orc_decoder* d);
uint32_t orc_dec_bins_ep(orc_decoder* d, int n);
unsigned orc_dec_trm(orc_decoder* d);
int orc_dec_finish(orc_decoder* d); /* 1 = both asserts of Decoder.cpp:75-81 hold */

/* ---- context initialisation (CABAC_ContextModelsInit.cpp:51-148) --------- */
void orc_map_prob_to_state(double p0, int* mps, int* state);
uint8_t orc_ctx_from_p0(double p0);
void orc_init_by_prob(const double* p0, int n, uint8_t* ctx);
void orc_init_by_state(const double* triples, int n, uint8_t* ctx);
uint8_t orc_matlab_uint8(double x); /* MATLAB uint8(): round half away, saturate (cabacEncode.m:30) */

/* ---- binarizer / debinarizer / finish detector --------------------------- */
int orc_binarize(uint32_t v, uint32_t Nq, int method, uint8_t* bins /* >= 80 */);
uint32_t orc_debinarize(const uint8_t* bins, int n, uint32_t Nq, int method);
/* cabacDecodeSymbolFinished.m:10-32; *n_p,*n_s as in the MATLAB (-1 / 0 at symbol start) */
int orc_symbol_finished(const uint8_t* g, int n, uint32_t Nq, int method, int* n_p, int* n_s);

/* ---- context selection ---------------------------------------------------- */
/* 0-based engine context index (i.e. MATLAB ctxID-1) or -1 for a bypass bin.
 * n is the 1-based bin position, g the symbol's own bins (only g[0..n-2] read),
 * up/up_len the bins of the up neighbour (ISS) or previous symbol (DEMO). */
int orc_select_ctx(int profile, int n, const uint8_t* g, const uint8_t* up, int up_len,
                   int Nlbp, unsigned types);
int orc_profile_num_ctx(int profile, int Nlbp);

/* ---- batch drivers over op arrays ---------------------------------------- */
int orc_encode_ops(uint32_t n_streams, const uint64_t* op_off, const void* ops, int op_width,
                   const uint8_t* ctx_init, uint32_t n_ctx, int per_stream_init,
                   uint8_t* out, uint64_t out_stride, uint32_t* out_len, int n_threads);
int orc_decode_ops(uint32_t n_streams, const uint64_t* byte_off, const uint8_t* bytes,
                   const uint64_t* op_off, const void* ops, int op_width,
                   const uint8_t* ctx_init, uint32_t n_ctx, int per_stream_init,
                   uint8_t* out_bins, uint8_t* finish_ok, int n_threads);

/* ---- symbol-level drivers (cabacEncode.m:45-70, cabacDecode.m:29-55,
 *      cabacDemo.m:101-129,143-186) ------------------------------------------
 * Stream s codes symbols[sym_off[s] .. sym_off[s+1]) in order.  For the ISS
 * profile the stream is a column-major matrix with `rows` rows: symbol i has an
 * up neighbour iff (i % rows) != 0 (cabacEncode.m:52).  For the DEMO profile the
 * "up" symbol is simply the previous symbol of the stream (cabacDemo.m:105). */
typedef struct {
  int profile, method;
  uint32_t Nq;
  int Nlbp;
  unsigned types;
  uint32_t rows; /* ISS only; 0 = whole stream is one column */
} orc_symcfg;
/* symbols -> ops (u8 op format); returns number of ops, writes at most cap */
uint64_t orc_symbols_to_ops(const orc_symcfg* cfg, const uint32_t* symbols, uint64_t n_sym,
                            uint8_t* ops, uint64_t cap);
int orc_encode_symbols(const orc_symcfg* cfg, uint32_t n_streams, const uint64_t* sym_off,
                       const uint32_t* symbols, const uint8_t* ctx_init, uint32_t n_ctx,
                       int per_stream_init, uint8_t* out, uint64_t out_stride, uint32_t* out_len,
                       uint32_t* bits_after_symbol /* optional, per symbol: getNumBits() after it */,
                       int n_threads);
int orc_decode_symbols(const orc_symcfg* cfg, uint32_t n_streams, const uint64_t* byte_off,
                       const uint8_t* bytes, const uint64_t* sym_off, const uint8_t* ctx_init,
                       uint32_t n_ctx, int per_stream_init, uint32_t* out_symbols,
                       uint8_t* finish_ok, int n_threads);

/* ---- ISS context-init statistics (cabacInitContextModel.m:15-129) -------- */
/* G: rows x cols column-major symbols; writes 7*Nlbp+2 probabilities p(0). */
void orc_iss_ctx_init(const uint32_t* G, uint32_t rows, uint32_t cols, uint32_t Nq, int method,
                      int Nlbp, unsigned types, double* p0_out);

#ifdef __cplusplus
}
#endif
#endif
