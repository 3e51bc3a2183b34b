// Library-level entry points: version, error string, launch counter.
#include "common.cuh"

#include <cstdlib>

namespace mg {
std::atomic<unsigned long long> g_launches{0};
static thread_local char t_error[512] = "";

bool pdl_enabled() {
    static int v = -1;
    if (v < 0) {
        const char* e = std::getenv("MAGGIE_B200_NO_PDL");
        v = (e && e[0] == '1') ? 0 : 1;
    }
    return v == 1;
}

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(t_error, sizeof(t_error), fmt, ap);
    va_end(ap);
}
}  // namespace mg

extern "C" {
int mg_version(void) { return 100; }  // 0.1.0
const char* mg_last_error(void) { return mg::t_error; }
unsigned long long mg_launch_count(void) { return mg::g_launches.load(); }
void mg_reset_launch_count(void) { mg::g_launches.store(0); }
}
