// One instantiation of the tensor-core ray-stage kernels per translation unit (-DNFB_RTC_INST=n).
#include "nfb_ray_tc.cuh"
#ifndef NFB_RTC_INST
#error "compile with -DNFB_RTC_INST=0..7"
#endif
#if NFB_RTC_INST == 0
int nfb_launch_ray_tc_fwd_p1(const nfbrtc::RayArgs& a, cudaStream_t st) { return nfbrtc::launch_ray_tc<1, false>(a, st); }
#elif NFB_RTC_INST == 1
int nfb_launch_ray_tc_fwd_p3(const nfbrtc::RayArgs& a, cudaStream_t st) { return nfbrtc::launch_ray_tc<3, false>(a, st); }
#elif NFB_RTC_INST == 2
int nfb_launch_ray_tc_bwd_p1(const nfbrtc::RayArgs& a, cudaStream_t st) { return nfbrtc::launch_ray_tc<1, true>(a, st); }
#elif NFB_RTC_INST == 3
int nfb_launch_ray_tc_bwd_p3(const nfbrtc::RayArgs& a, cudaStream_t st) { return nfbrtc::launch_ray_tc<3, true>(a, st); }
#elif NFB_RTC_INST == 4
int nfb_launch_ray_tc_fwd_p1_save(const nfbrtc::RayArgs& a, cudaStream_t st) { return nfbrtc::launch_ray_tc<1, false, true>(a, st); }
#elif NFB_RTC_INST == 5
int nfb_launch_ray_tc_fwd_p3_save(const nfbrtc::RayArgs& a, cudaStream_t st) { return nfbrtc::launch_ray_tc<3, false, true>(a, st); }
#elif NFB_RTC_INST == 6
int nfb_launch_ray_tc_bwd_stash_p1(const nfbrtc::RayArgs& a, cudaStream_t st) { return nfbrtc::launch_ray_tc_bwd_stash<1>(a, st); }
#elif NFB_RTC_INST == 7
int nfb_launch_ray_tc_bwd_stash_p3(const nfbrtc::RayArgs& a, cudaStream_t st) { return nfbrtc::launch_ray_tc_bwd_stash<3>(a, st); }
#endif
