// One instantiation of the tensor-core ray-stage kernels per translation unit (-DNFB_RTC_INST=n).
#include "nfb_ray_tc.cuh"
#ifndef NFB_RTC_INST
#error "compile with -DNFB_RTC_INST=0..3"
#endif
#if NFB_RTC_INST == 0
int nfb_launch_ray_tc_fwd_p1(const nfbrtc::RayArgs& a, cudaStream_t st) { return nfbrtc::launch_ray_tc<1, false>(a, st); }
#elif NFB_RTC_INST == 1
int nfb_launch_ray_tc_fwd_p3(const nfbrtc::RayArgs& a, cudaStream_t st) { return nfbrtc::launch_ray_tc<3, false>(a, st); }
#elif NFB_RTC_INST == 2
int nfb_launch_ray_tc_bwd_p1(const nfbrtc::RayArgs& a, cudaStream_t st) { return nfbrtc::launch_ray_tc<1, true>(a, st); }
#elif NFB_RTC_INST == 3
int nfb_launch_ray_tc_bwd_p3(const nfbrtc::RayArgs& a, cudaStream_t st) { return nfbrtc::launch_ray_tc<3, true>(a, st); }
#endif
