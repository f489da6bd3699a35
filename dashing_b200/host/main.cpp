// dashing_b200 — thin CLI over the host layer: `dashing_b200 sketch ...` / `dashing_b200 dist|cmp ...` with the hot
// subset of dashing's flags (src/dashing.cpp:294-409, src/distmain.cpp:28-204).
#include <cstdio>
#include <unistd.h>
extern "C" int db200h_cli(int argc, char **argv);
int main(int argc, char **argv) {
    const int rc = db200h_cli(argc, argv);
    // every output file has been closed by now; leave without the CUDA runtime's atexit teardown (about a second on a
    // B200 node) — the kernel driver reclaims the context either way
    std::fflush(nullptr);
    _exit(rc);
}
