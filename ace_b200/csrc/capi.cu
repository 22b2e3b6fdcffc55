// Library-level C ABI: errors, options, launch counter, GEMM dispatcher, development GEMM hook.
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include <nvtx3/nvToolsExt.h>  // header-only NVTX v3: the injection library is looked up at run time, nothing to link

#include "gemm.cuh"
#include "kernels.cuh"

namespace ace {

std::atomic<long long> g_launch_count{0};
std::atomic<long long> g_umma_count{0}, g_simt_count{0};

static thread_local std::string t_last_error;
void set_last_error(const char* msg) { t_last_error = msg ? msg : ""; }

Options& options() {
  static Options o = [] {
    Options x;
    if (const char* e = getenv("ACE_B200_PDL")) x.pdl = atoi(e);
    // every GEMM on the SIMT kernel: for compute-sanitizer racecheck / synccheck runs (tools/sanitize.sh), which do not model
    // the asynchronous tcgen05 / TMA proxies
    if (const char* e = getenv("ACE_B200_FORCE_SIMT")) x.force_simt = atoi(e) ? 1 : 0;
    if (const char* e = getenv("ACE_B200_NVTX")) x.nvtx = atoi(e) ? 1 : 0;
    if (const char* e = getenv("ACE_B200_CLN_GEMM")) x.cln_gemm = atoi(e) ? 1 : 0;
    return x;
  }();
  return o;
}

namespace {
struct ProfRec {
  std::string name;
  cudaEvent_t start, stop;
};
std::vector<ProfRec> g_prof;
}  // namespace

namespace {
ace_scope_callback g_scope_cb = nullptr;
void* g_scope_user = nullptr;
}  // namespace

ProfileScope::ProfileScope(const char* n, cudaStream_t s) : name(n), stream(s) {
  if (g_scope_cb) {
    g_scope_cb(n, 1, g_scope_user);
    hooked = true;
  }
  if (options().nvtx) {  // the reference names its timed regions (Timer.child, fme/core/benchmark/timer.py); here: one range per operator
    nvtxRangePushA(n);
    ranged = true;
  }
  if (!options().profile) return;
  if (cudaEventCreate(&start) != cudaSuccess || cudaEventCreate(&stop) != cudaSuccess) {
    start = stop = nullptr;
    return;
  }
  cudaEventRecord(start, stream);
}
ProfileScope::~ProfileScope() {
  if (ranged) nvtxRangePop();
  if (start) {
    cudaEventRecord(stop, stream);
    g_prof.push_back({name, start, stop});
  }
  if (hooked && g_scope_cb) g_scope_cb(name, 0, g_scope_user);
}

void run_gemm(const GemmOp& op, cudaStream_t stream) {
  ProfileScope prof(op.name, stream);
  const char* why = nullptr;
  if (!options().force_simt && umma_eligible(op, &why)) {
    run_gemm_umma(op, stream);
  } else if (op.bfly && !options().force_simt && umma_eligible(bfly_dense(op), &why)) {
    run_gemm_umma(bfly_dense(op), stream);  // the butterfly variant is not compiled for this shape: same result, twice the MMAs
  } else {
    run_gemm_simt(op, stream);
  }
}

}  // namespace ace

using namespace ace;

extern "C" int ace_version(void) { return 100; }

extern "C" const char* ace_last_error(void) { return t_last_error.c_str(); }

extern "C" int ace_set_option(const char* key, int value) {
  ACE_API_BEGIN
  ACE_REQUIRE(key != nullptr, "ace_set_option: null key");
  if (!strcmp(key, "force_simt")) options().force_simt = value;
  else if (!strcmp(key, "profile")) options().profile = value;
  else if (!strcmp(key, "nvtx")) options().nvtx = value ? 1 : 0;
  else if (!strcmp(key, "split_terms")) {
    ACE_REQUIRE(value == 1 || value == 3, "split_terms must be 1 or 3");
    options().split_terms = value;
  } else if (!strcmp(key, "pair")) {
    options().pair = value < 0 ? -1 : (value ? 1 : 0);
  } else if (!strcmp(key, "conv_bn")) {
    ACE_REQUIRE(value == 0 || value == 192 || value == 256, "conv_bn must be 0, 192 or 256");
    options().conv_bn = value;
  } else if (!strcmp(key, "pdl")) {
    options().pdl = value;
  } else if (!strcmp(key, "dbg")) {
    options().dbg = value;
  } else if (!strcmp(key, "l2_persist")) {
    options().l2_persist = value ? 1 : 0;
  } else if (!strcmp(key, "inv2")) {
    options().inv2 = value ? 1 : 0;
  } else if (!strcmp(key, "tile_list")) {
    options().tile_list = value ? 1 : 0;
  } else if (!strcmp(key, "dhconv_t")) {
    options().dhconv_t = value ? 1 : 0;
  } else if (!strcmp(key, "group_order")) {
    options().group_order = value ? 1 : 0;
  } else if (!strcmp(key, "mma_batch")) {
    options().mma_batch = value ? 1 : 0;
  } else if (!strcmp(key, "sp")) {
    options().sp = value;  // 0 off, 1 where it pays, 2 wherever eligible
  } else if (!strcmp(key, "sp_tma")) {
    options().sp_tma = value ? 1 : 0;
  } else if (!strcmp(key, "bfly_pair")) {
    options().bfly_pair = value ? 1 : 0;
  } else if (!strcmp(key, "tile_serpentine")) {
    options().tile_serpentine = value ? 1 : 0;
  } else if (!strcmp(key, "sp_tmx")) {
    options().sp_tmx = value ? 1 : 0;
  } else if (!strcmp(key, "cln_gemm")) {
    options().cln_gemm = value ? 1 : 0;
  } else if (!strcmp(key, "trace")) {
    options().trace = value ? 1 : 0;
  } else if (!strcmp(key, "umma_bk")) {
    ACE_REQUIRE(value == 0 || value == 32 || value == 64, "umma_bk must be 0, 32 or 64");
    options().umma_bk = value;
  } else if (!strcmp(key, "umma_bn")) {
    ACE_REQUIRE(value == 0 || value == 128 || value == 192 || value == 256, "umma_bn must be 0, 128, 192 or 256");
    options().umma_bn = value;
  } else ACE_REQUIRE(false, "ace_set_option: unknown option '%s'", key);
  ACE_API_END
}

extern "C" int ace_get_option(const char* key) {
  if (!key) return -1;
  if (!strcmp(key, "force_simt")) return options().force_simt;
  if (!strcmp(key, "profile")) return options().profile;
  if (!strcmp(key, "nvtx")) return options().nvtx;
  if (!strcmp(key, "split_terms")) return options().split_terms;
  if (!strcmp(key, "count_umma")) return (int)g_umma_count.load();
  if (!strcmp(key, "count_simt")) return (int)g_simt_count.load();
  if (!strcmp(key, "umma_bn")) return options().umma_bn;
  if (!strcmp(key, "pair")) return options().pair;
  if (!strcmp(key, "umma_bk")) return options().umma_bk;
  if (!strcmp(key, "dhconv_t")) return options().dhconv_t;
  if (!strcmp(key, "tile_list")) return options().tile_list;
  if (!strcmp(key, "inv2")) return options().inv2;
  if (!strcmp(key, "sp")) return options().sp;
  if (!strcmp(key, "sp_tma")) return options().sp_tma;
  if (!strcmp(key, "sp_tmx")) return options().sp_tmx;
  if (!strcmp(key, "mma_batch")) return options().mma_batch;
  if (!strcmp(key, "bfly_pair")) return options().bfly_pair;
  if (!strcmp(key, "group_order")) return options().group_order;
  if (!strcmp(key, "tile_serpentine")) return options().tile_serpentine;
  if (!strcmp(key, "cln_gemm")) return options().cln_gemm;
  if (!strcmp(key, "pdl")) return options().pdl;
  if (!strcmp(key, "trace")) return options().trace;
  if (!strcmp(key, "conv_bn")) return options().conv_bn;
  if (!strcmp(key, "dbg")) return options().dbg;
  return -1;
}

extern "C" int ace_set_scope_callback(ace_scope_callback cb, void* user) {
  g_scope_cb = cb;
  g_scope_user = user;
  return ACE_OK;
}

extern "C" int ace_debug_scope(const char* name) {
  ACE_API_BEGIN
  ACE_REQUIRE(name != nullptr, "ace_debug_scope: null name");
  static std::string held;  // the scope keeps the pointer until it closes
  held = name;
  { ProfileScope scope(held.c_str(), nullptr); }
  ACE_API_END
}

extern "C" long long ace_launch_count(void) { return g_launch_count.load(); }

// Synchronises the device, writes "name count total_ms\n" lines (aggregated, launch order of first
// appearance) into buf and clears the records.  Returns the number of bytes written (excluding NUL).
extern "C" int ace_profile_report(char* buf, int buflen) {
  if (!buf || buflen <= 0) return 0;
  cudaDeviceSynchronize();
  std::vector<std::string> order;
  std::vector<double> total;
  std::vector<long long> count;
  for (auto& r : g_prof) {
    float ms = 0.f;
    cudaEventElapsedTime(&ms, r.start, r.stop);
    cudaEventDestroy(r.start);
    cudaEventDestroy(r.stop);
    size_t i = 0;
    for (; i < order.size(); ++i)
      if (order[i] == r.name) break;
    if (i == order.size()) {
      order.push_back(r.name);
      total.push_back(0.0);
      count.push_back(0);
    }
    total[i] += ms;
    count[i] += 1;
  }
  g_prof.clear();
  std::string out;
  for (size_t i = 0; i < order.size(); ++i) out += strprintf("%s %lld %.6f\n", order[i].c_str(), count[i], total[i]);
  int n = (int)std::min<size_t>(out.size(), (size_t)buflen - 1);
  memcpy(buf, out.data(), n);
  buf[n] = 0;
  return n;
}

// D[z][m][n] = sum_k A[z][m][k] * B[z][n][k] through the split-plane machinery (tests only).
// layout bit 0: A is stored MN-major ([z][k][m]); bit 1: B is stored MN-major ([z][k][n]).
extern "C" int ace_dev_gemm(const float* a_dev, const float* b_dev, float* d_dev, int m, int n, int k, int nbatch,
                            int layout, int impl, void* stream) {
  ACE_API_BEGIN
  ACE_REQUIRE(a_dev && b_dev && d_dev && m > 0 && n > 0 && k > 0 && nbatch > 0, "ace_dev_gemm: bad argument");
  ACE_REQUIRE(layout >= 0 && layout <= 3, "ace_dev_gemm: layout must be 0..3");
  cudaStream_t s = (cudaStream_t)stream;
  const bool a_mn = layout & 1, b_mn = (layout & 2) != 0;
  const int kp = (int)round_up(k, 8), mp = (int)round_up(m, 8), np = (int)round_up(n, 8);
  const long long a_per = a_mn ? (long long)k * mp : (long long)m * kp;
  const long long b_per = b_mn ? (long long)k * np : (long long)n * kp;
  DevBuf abuf, bbuf;
  abuf.ensure(2 * (size_t)(a_per * nbatch) * sizeof(bf16));
  bbuf.ensure(2 * (size_t)(b_per * nbatch) * sizeof(bf16));
  if (a_mn) launch_split_pad(a_dev, (long long)nbatch * k, m, mp, abuf.as<bf16>(), a_per * nbatch, s);
  else launch_split_pad(a_dev, (long long)nbatch * m, k, kp, abuf.as<bf16>(), a_per * nbatch, s);
  if (b_mn) launch_split_pad(b_dev, (long long)nbatch * k, n, np, bbuf.as<bf16>(), b_per * nbatch, s);
  else launch_split_pad(b_dev, (long long)nbatch * n, k, kp, bbuf.as<bf16>(), b_per * nbatch, s);
  GemmOp op = make_gemm_op("dev_gemm");
  op.M = m;
  op.N = n;
  op.K = k;
  op.Z2 = nbatch;
  if (a_mn) op.A = {abuf.as<bf16>(), a_per * nbatch, 1, (long long)mp, 0, a_per};
  else op.A = {abuf.as<bf16>(), a_per * nbatch, (long long)kp, 1, 0, a_per};
  if (b_mn) op.B = {bbuf.as<bf16>(), b_per * nbatch, 1, (long long)np, 0, b_per};
  else op.B = {bbuf.as<bf16>(), b_per * nbatch, (long long)kp, 1, 0, b_per};
  op.epi.flags = EPI_OUT_F32;
  op.epi.outf = d_dev;
  op.epi.f_z2 = (long long)m * n;
  op.epi.f_m0 = n;
  op.epi.f_n = 1;
  if (impl == 0) {
    run_gemm_simt(op, s);
  } else {
    const char* why = nullptr;
    if (!umma_eligible(op, &why)) throw Error(ACE_ERR_INVALID, std::string("ace_dev_gemm: not eligible for tcgen05 kernel: ") + (why ? why : "?"));
    run_gemm_umma(op, s);
  }
  ACE_CHECK_CUDA(cudaStreamSynchronize(s));  // temporaries are freed on return
  ACE_API_END
}
