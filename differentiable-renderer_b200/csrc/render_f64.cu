// render_f64.cu — the analytic-scene render kernels in IEEE double: the parity instantiation
// (the reference computes in double, src/render.cpp:22).  See render_kernels.cuh.
#include "render_kernels.cuh"

namespace drtbh {

int launch_analytic_f64(drtb_ctx* ctx, drtb::RenderArgs& a, const AnalyticLaunch& l, cudaStream_t stream, size_t& rows)
{
    return drtb::launch_analytic<double>(ctx, ctx->sc64, a, l, stream, rows);
}

cudaError_t init_tables_render_f64() { return drtb::upload_sincos_tab(); }

} // namespace drtbh
