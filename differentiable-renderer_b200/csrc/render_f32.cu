// render_f32.cu — the analytic-scene render kernels in float: the throughput instantiation (image and
// gradient accumulators stay double).  See render_kernels.cuh.
#include "render_kernels.cuh"

namespace drtbh {

int launch_analytic_f32(drtb_ctx* ctx, drtb::RenderArgs& a, const AnalyticLaunch& l, cudaStream_t stream, size_t& rows)
{
    return drtb::launch_analytic<float>(ctx, ctx->sc32, a, l, stream, rows);
}

} // namespace drtbh
