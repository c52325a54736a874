// Library-wide state: version, thread-local error text, launch counter.
#include "common.cuh"

namespace tagrec {
thread_local std::string g_last_error;
std::atomic<uint64_t> g_launches{0};
}  // namespace tagrec

extern "C" int tagrec_version(void) { return 100; }
extern "C" const char* tagrec_last_error(void) { return tagrec::g_last_error.c_str(); }
extern "C" uint64_t tagrec_launch_count(void) { return tagrec::g_launches.load(); }
extern "C" size_t tagrec_sizeof_struct(int which) {
    switch (which) {
        case 0: return sizeof(tagrec_csr_t);
        case 1: return sizeof(tagrec_mirror_t);
        case 2: return sizeof(tagrec_route_plan_t);
        case 3: return sizeof(tagrec_adam_t);
        default: return 0;
    }
}
