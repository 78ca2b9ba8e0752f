// Library bookkeeping for the C ABI declared in include/spateo_b200.h.
#include <atomic>

#include "common.cuh"

static std::atomic<int64_t> g_launches{0};

void spb_count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

extern "C" int64_t spb_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

extern "C" int spb_version(void) { return 100; }

extern "C" int spb_sizeof_em_params(void) { return (int)sizeof(spb_em_params); }
extern "C" int spb_sizeof_scalars(void) { return (int)sizeof(spb_scalars); }
extern "C" int spb_sizeof_field_desc(void) { return (int)sizeof(spb_field_desc); }
