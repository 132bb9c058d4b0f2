// One instantiation of the tensor-core view-stage backward kernels per translation unit (-DNFB_VTCB_INST=n).
#include "nfb_view_tc_bwd.cuh"
#ifndef NFB_VTCB_INST
#error "compile with -DNFB_VTCB_INST=0..3"
#endif
#if NFB_VTCB_INST == 0
int nfb_launch_view_tc_bwd_p1_fused(const nfbview::ViewArgs& a, cudaStream_t st) { return nfbvtcb::launch_view_tc_bwd<1, true>(a, st); }
#elif NFB_VTCB_INST == 1
int nfb_launch_view_tc_bwd_p3_fused(const nfbview::ViewArgs& a, cudaStream_t st) { return nfbvtcb::launch_view_tc_bwd<3, true>(a, st); }
#elif NFB_VTCB_INST == 2
int nfb_launch_view_tc_bwd_p1_tensor(const nfbview::ViewArgs& a, cudaStream_t st) { return nfbvtcb::launch_view_tc_bwd<1, false>(a, st); }
#elif NFB_VTCB_INST == 3
int nfb_launch_view_tc_bwd_p3_tensor(const nfbview::ViewArgs& a, cudaStream_t st) { return nfbvtcb::launch_view_tc_bwd<3, false>(a, st); }
#endif
