// Stand-in for the CMake-generated scalable_ccd/config.hpp
// (reference: src/scalable_ccd/config.hpp.in:1-14).  TEST INFRASTRUCTURE ONLY.
#pragma once
#define SCALABLE_CCD_NAME "scalable_ccd"
#define SCALABLE_CCD_VER "0.1.0"
#define SCALABLE_CCD_VER_MAJOR "0"
#define SCALABLE_CCD_VER_MINOR "1"
#define SCALABLE_CCD_VER_PATCH "0"
#ifdef REF_WITH_CUDA
#define SCALABLE_CCD_WITH_CUDA
#endif
#ifndef REF_USE_FLOAT /* float build: scalar.hpp:16-18 */
#define SCALABLE_CCD_USE_DOUBLE
#endif
#ifdef REF_TOI_PER_QUERY
#define SCALABLE_CCD_TOI_PER_QUERY
#endif
// SCALABLE_CCD_WITH_PROFILER left undefined (would need nlohmann/json)
