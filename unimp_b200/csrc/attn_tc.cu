// placeholder — replaced by the tcgen05 implementation
#include "common.cuh"
namespace unimp {
bool attn_fwd_tc_supported(unimp_view_t, unimp_view_t, unimp_view_t, unimp_mview_t, const int32_t*, int, int, int, int) { return false; }
int launch_attn_fwd_tc(unimp_view_t, unimp_view_t, unimp_view_t, const int32_t*, unimp_mview_t, float*, int, int, int, int, int, int, float, cudaStream_t) { set_error("attn_fwd_tc: not built"); return UNIMP_E_SHAPE; }
bool attn_bwd_tc_supported(unimp_view_t, unimp_view_t, unimp_view_t, unimp_view_t, unimp_mview_t, unimp_mview_t, unimp_mview_t, const int32_t*, int, int, int, int) { return false; }
int launch_attn_bwd_tc(unimp_view_t, unimp_view_t, unimp_view_t, const int32_t*, unimp_view_t, unimp_view_t, const float*, void*, unimp_mview_t, unimp_mview_t, unimp_mview_t, int, int, int, int, int, int, float, cudaStream_t) { set_error("attn_bwd_tc: not built"); return UNIMP_E_SHAPE; }
}
