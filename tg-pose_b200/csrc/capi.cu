// capi.cu -- library-wide state and the dispatching entry points of libtgpose_b200.so.
#include "common.cuh"

namespace tgp {
thread_local char g_err[256] = "";
std::atomic<unsigned long long> g_launches{0};
}  // namespace tgp

using namespace tgp;

int tgp_gemm_validate(const tgp_gemm_args* a);
int tgp_gemm_simt(const tgp_gemm_args* a, cudaStream_t st);
int tgp_gemm_tc(const tgp_gemm_args* a, cudaStream_t st);

extern "C" int tgp_version(void) { return 100; }  // 0.1.0

extern "C" const char* tgp_last_error(void) { return g_err; }

extern "C" unsigned long long tgp_launch_count(void) { return g_launches.load(); }

extern "C" int tgp_gemm(const tgp_gemm_args* args_host, tgp_stream_t stream) {
    int rc = tgp_gemm_validate(args_host);
    if (rc) return rc;
    if (args_host->A_split && args_host->B_split) return tgp_gemm_tc(args_host, as_stream(stream));
    if (!args_host->A || !args_host->Bmat) return fail(TGP_EINVAL, "tgp_gemm: null operand");
    return tgp_gemm_simt(args_host, as_stream(stream));
}
