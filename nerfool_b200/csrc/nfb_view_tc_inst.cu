// One instantiation of the tensor-core view-stage kernels per translation unit (-DNFB_VTC_INST=n).
#include "nfb_view_tc.cuh"
#ifndef NFB_VTC_INST
#error "compile with -DNFB_VTC_INST=0..3"
#endif
#if NFB_VTC_INST == 0
int nfb_launch_view_tc_fwd_p1_fused(const nfbview::ViewArgs& a, cudaStream_t st) { return nfbvtc::launch_view_tc_fwd<1, true>(a, st); }
#elif NFB_VTC_INST == 1
int nfb_launch_view_tc_fwd_p3_fused(const nfbview::ViewArgs& a, cudaStream_t st) { return nfbvtc::launch_view_tc_fwd<3, true>(a, st); }
#elif NFB_VTC_INST == 2
int nfb_launch_view_tc_fwd_p1_tensor(const nfbview::ViewArgs& a, cudaStream_t st) { return nfbvtc::launch_view_tc_fwd<1, false>(a, st); }
#elif NFB_VTC_INST == 3
int nfb_launch_view_tc_fwd_p3_tensor(const nfbview::ViewArgs& a, cudaStream_t st) { return nfbvtc::launch_view_tc_fwd<3, false>(a, st); }
#endif
