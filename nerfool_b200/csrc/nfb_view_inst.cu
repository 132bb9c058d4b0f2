// One instantiation of the view-stage kernel per translation unit, selected with -DNFB_VIEW_INST=n
// (keeps each nvcc invocation short and lets the build run them in parallel).
#include "nfb_view_stage.cuh"
#ifndef NFB_VIEW_INST
#error "compile with -DNFB_VIEW_INST=0..5"
#endif
#if NFB_VIEW_INST == 0
int nfb_launch_view_tensor_fwd(const nfbview::ViewArgs& a, cudaStream_t st) {
  return nfbview::launch_view<false, false, 4>(a, st, "k_view_stage<tensor,fwd>");
}
#elif NFB_VIEW_INST == 1
int nfb_launch_view_fused_fwd(const nfbview::ViewArgs& a, cudaStream_t st) {
  return nfbview::launch_view<true, false, 4>(a, st, "k_view_stage<fused,fwd>");
}
#elif NFB_VIEW_INST == 2
int nfb_launch_view_tensor_bwd(const nfbview::ViewArgs& a, cudaStream_t st) {
  return nfbview::launch_view<false, true, 4>(a, st, "k_view_stage<tensor,bwd>");
}
#elif NFB_VIEW_INST == 3
int nfb_launch_view_fused_bwd(const nfbview::ViewArgs& a, cudaStream_t st) {
  return nfbview::launch_view<true, true, 4>(a, st, "k_view_stage<fused,bwd>");
}
#elif NFB_VIEW_INST == 4
int nfb_launch_view_tensor_wgrad(const nfbview::ViewArgs& a, cudaStream_t st) {
  return nfbview::launch_view<false, true, 2, true>(a, st, "k_view_stage<tensor,bwd,wgrad>");
}
#elif NFB_VIEW_INST == 5
int nfb_launch_view_fused_wgrad(const nfbview::ViewArgs& a, cudaStream_t st) {
  return nfbview::launch_view<true, true, 2, true>(a, st, "k_view_stage<fused,bwd,wgrad>");
}
#endif
