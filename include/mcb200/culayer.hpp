// mcb200/culayer.hpp -- the reference's existing GPU operator boundary,
// include/culayer/culayer.hpp:6-13, same C++ signature (so the mangled symbol
// _Z10cusimulateiP12particle_tagPKfS2_Pfiif is the one src/layer.cpp:259 and
// src/test_culayer.cu:69 link against), implemented on the B200 path.
//
// Contract kept: host pointers, caller-owned arrays, blocking; the tally is
// in/out (the absorbed weight is ADDED to weights_absorbed); on return every
// particle lies outside [min_index, max_index) and `particles` holds all n
// final states (in unspecified order -- the reference's own order is
// scrambled by its host-side sort, src/culayer.cu:82-86).  Like the
// reference kernel there is no particle_min_weight cut-off
// (src/culayer_kernel.cu:53).  Errors print and exit like gpu_errcheck.
#ifndef MCB200_CULAYER_HPP
#define MCB200_CULAYER_HPP

#include "layer.hpp"

void cusimulate(int n, Particle *particles, float const *const sigs,
                float const *const absorption_rates, float *const weights_absorbed,
                int min_index, int max_index, float dx);

#endif
