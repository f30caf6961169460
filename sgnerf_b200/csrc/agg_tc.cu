// agg_tc.cu -- bf16 tensor-core (tcgen05/TMEM/TMA) aggregator path.  Placeholder until the fused kernel lands.
#include "agg_common.cuh"

using namespace sgn;

int sgn_agg_tc_workspace_bytes(const AggPlan&, int64_t, int, int, size_t*)
{
    set_error("aggregator: SGN_PRECISION_BF16 is not built in this revision");
    return SGN_E_INVALID;
}

int sgn_agg_tc_forward(const AggPlan&, const float* const*, const float* const*, const SgnPointTables*, const int32_t*, const float*,
                       const float*, const float*, const float*, int64_t, int, int, float*, uint8_t*, float*, float*, float*, void*, size_t,
                       cudaStream_t)
{
    set_error("aggregator: SGN_PRECISION_BF16 is not built in this revision");
    return SGN_E_INVALID;
}
