// Instantiations of the stash-based tensor-core view-stage kernels (-DNFB_VTCS_INST=n): the forward variants that
// write the activation stash and the backward kernels that read it.
#ifndef NFB_VTCS_INST
#error "compile with -DNFB_VTCS_INST=0..3"
#endif
#if NFB_VTCS_INST == 0
#include "nfb_view_tc.cuh"
int nfb_launch_view_tc_fwd_p1_fused_save(const nfbview::ViewArgs& a, cudaStream_t st) { return nfbvtc::launch_view_tc_fwd<1, true, true>(a, st); }
#elif NFB_VTCS_INST == 1
#include "nfb_view_tc.cuh"
int nfb_launch_view_tc_fwd_p3_fused_save(const nfbview::ViewArgs& a, cudaStream_t st) { return nfbvtc::launch_view_tc_fwd<3, true, true>(a, st); }
#elif NFB_VTCS_INST == 2
#include "nfb_view_tc_bwd2.cuh"
int nfb_launch_view_tc_bwd_stash_p1(const nfbview::ViewArgs& a, cudaStream_t st) { return nfbvtcs::launch_view_tc_bwd_stash<1>(a, st); }
#elif NFB_VTCS_INST == 3
#include "nfb_view_tc_bwd2.cuh"
int nfb_launch_view_tc_bwd_stash_p3(const nfbview::ViewArgs& a, cudaStream_t st) { return nfbvtcs::launch_view_tc_bwd_stash<3>(a, st); }
#endif
