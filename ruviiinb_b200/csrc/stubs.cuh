// temporary: entry points not built yet
extern "C" {
int ruvnb_disp_newton(ruvnb_ctx* ctx, const ruvnb_opts*, double*, int) { return ctx ? ctx->fail(RUVNB_ESTATE, "not built yet") : RUVNB_EARG; }
}
