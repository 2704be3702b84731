// Error reporting for the C ABI (thread-local last-error string) and version query.
#include "common.cuh"
#include "../../include/margipose_b200.h"
#include <stdarg.h>

static thread_local char g_err[512] = "";

void mp_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

extern "C" {
int mp_abi_version(void) { return 4; }   // 4: split (bf16x3) mode (lo_delta, acc_in), epilogue affine, mp_bn_fold_eval, mp_stem_im2col_u8, hyper[5]
const char* mp_last_error(void) { return g_err; }
}

// ---- tunables (experiments only; defaults are what the benchmarks use)
#include <string.h>
void mp_set_igemm_smem(long long v);
void mp_set_igemm_split_n(long long v);
void mp_set_igemm_halo(long long v);
void mp_set_igemm_pair(long long v);
void mp_set_igemm_dbg(long long v);
void mp_set_igemm_prefetch_b(long long v);
void mp_set_igemm_resident(long long v);
void mp_set_igemm_astages(long long v);
void mp_set_igemm_trace(long long v);
void mp_set_igemm_mt(int which, long long v);
void mp_set_wgrad_tunable(int which, long long v);
extern long long g_tail_fast;
extern long long g_tail_waves;
extern long long g_tail_wpj;
extern long long g_tail_cap;
extern long long g_tail_wpj_max;
extern long long g_bn_tma;

static long long g_pdl = 1;
int mp_pdl_enabled() { return g_pdl != 0; }
static long long g_deterministic = 0;
int mp_deterministic() { return g_deterministic != 0; }

extern "C" int mp_set_tunable(const char* name, int64_t value) {
  if (!name) { mp_set_error("mp_set_tunable: null name"); return MP_ERR_ARG; }
  if (!strcmp(name, "pdl")) { g_pdl = value; return MP_OK; }
  if (!strcmp(name, "deterministic")) { g_deterministic = value; return MP_OK; }
  if (!strcmp(name, "igemm_smem")) { mp_set_igemm_smem(value); return MP_OK; }
  if (!strcmp(name, "igemm_split_n")) { mp_set_igemm_split_n(value); return MP_OK; }
  if (!strcmp(name, "igemm_halo")) { mp_set_igemm_halo(value); return MP_OK; }
  if (!strcmp(name, "igemm_pair")) { mp_set_igemm_pair(value); return MP_OK; }
  if (!strcmp(name, "igemm_dbg")) { mp_set_igemm_dbg(value); return MP_OK; }
  if (!strcmp(name, "igemm_prefetch_b")) { mp_set_igemm_prefetch_b(value); return MP_OK; }
  if (!strcmp(name, "igemm_resident")) { mp_set_igemm_resident(value); return MP_OK; }
  if (!strcmp(name, "igemm_astages")) { mp_set_igemm_astages(value); return MP_OK; }
  if (!strcmp(name, "igemm_trace")) { mp_set_igemm_trace(value); return MP_OK; }
  if (!strcmp(name, "igemm_mt")) { mp_set_igemm_mt(0, value); return MP_OK; }
  if (!strcmp(name, "igemm_mt_ctas")) { mp_set_igemm_mt(1, value); return MP_OK; }
  if (!strcmp(name, "igemm_ctas")) { mp_set_igemm_mt(2, value); return MP_OK; }
  if (!strcmp(name, "wgrad_ctas")) { mp_set_wgrad_tunable(0, value); return MP_OK; }
  if (!strcmp(name, "wgrad_halo")) { mp_set_wgrad_tunable(1, value); return MP_OK; }
  if (!strcmp(name, "wgrad_dbg")) { mp_set_wgrad_tunable(2, value); return MP_OK; }
  if (!strcmp(name, "wgrad_kp")) { mp_set_wgrad_tunable(3, value); return MP_OK; }
  if (!strcmp(name, "wgrad_slice")) { mp_set_wgrad_tunable(4, value); return MP_OK; }
  if (!strcmp(name, "wgrad_smem")) { mp_set_wgrad_tunable(5, value); return MP_OK; }
  if (!strcmp(name, "tail_fast")) { g_tail_fast = value; return MP_OK; }
  if (!strcmp(name, "tail_ctas_per_sm")) { g_tail_waves = value; return MP_OK; }
  if (!strcmp(name, "tail_wpj")) { g_tail_wpj = value; return MP_OK; }
  if (!strcmp(name, "tail_wpj_max")) { g_tail_wpj_max = value; return MP_OK; }
  if (!strcmp(name, "tail_cap")) { g_tail_cap = value == 8 ? 8 : 4; return MP_OK; }
  if (!strcmp(name, "bn_tma")) { g_bn_tma = value; return MP_OK; }
  mp_set_error("mp_set_tunable: unknown tunable '%s'", name);
  return MP_ERR_ARG;
}
