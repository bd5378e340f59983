// render_f64.cu — the analytic-scene render kernels in IEEE double: the parity instantiation
// (the reference computes in double, src/render.cpp:22).  See render_kernels.cuh.
#include "render_kernels.cuh"

namespace drtbh {

int launch_analytic_f64(drtb_ctx* ctx, drtb::RenderArgs& a, const AnalyticLaunch& l, cudaStream_t stream, size_t& rows)
{
    return drtb::launch_analytic<double>(ctx, ctx->sc64, a, l, stream, rows);
}

int launch_retrace_f64(drtb_ctx* ctx, drtb::RenderArgs& a, int P3, bool want_grad, size_t row0, cudaStream_t stream, size_t& rows_added)
{
    return drtb::launch_retrace<double>(ctx, ctx->sc64, a, P3, want_grad, row0, stream, rows_added);
}

cudaError_t init_tables_render_f64() { return drtb::upload_sincos_tab(); }

} // namespace drtbh
