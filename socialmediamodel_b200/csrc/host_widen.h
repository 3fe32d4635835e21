// host_widen.h — host thread pool that widens FP32 ensemble outputs to Float64 (see host_widen.cpp).
#pragma once
#include <cstddef>

struct smm_hostpool;
smm_hostpool* smm_hostpool_create(int nthreads);  // nthreads <= 0: default (cores / local ranks, <= 16)
void smm_hostpool_destroy(smm_hostpool* p);
int smm_hostpool_threads(const smm_hostpool* p);
// dst[i] = (double)src[i], i < count; returns when all slices are written (the caller's thread works too)
void smm_hostpool_widen(smm_hostpool* p, const float* src, double* dst, size_t count);
