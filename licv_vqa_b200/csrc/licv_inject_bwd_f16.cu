#include "licv_inject_impl.cuh"
namespace licv { namespace inject { LICV_DEFINE_RUN_BWD(LICV_F16) } }
#ifdef LICV_TRACE
extern "C" int licv_debug_read_pipe_trace(unsigned long long* host, int n) {
    return (int)cudaMemcpyFromSymbol(host, licv::inject::g_pipe_trace, sizeof(unsigned long long) * n);
}
#endif
