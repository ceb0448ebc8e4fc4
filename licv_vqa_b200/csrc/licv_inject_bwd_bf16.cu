#include "licv_inject_impl.cuh"
namespace licv { namespace inject { LICV_DEFINE_RUN_BWD(LICV_BF16) } }
