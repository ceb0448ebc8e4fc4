#include "licv_inject_impl.cuh"
namespace licv { namespace inject { LICV_DEFINE_RUN_FWD(LICV_F16) } }
