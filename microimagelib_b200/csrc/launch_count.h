// Process-wide count of kernels launched by this library (bench.py reports it as gpu_launches).
#pragma once
#include <atomic>
extern std::atomic<long long> g_milb_launches;
static inline void milb_count_launches(int n) { g_milb_launches.fetch_add(n, std::memory_order_relaxed); }
