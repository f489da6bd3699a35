// dashing_b200 — thin CLI over the host layer: `dashing_b200 sketch ...` / `dashing_b200 dist|cmp ...` with the hot
// subset of dashing's flags (src/dashing.cpp:294-409, src/distmain.cpp:28-204).
extern "C" int db200h_cli(int argc, char **argv);
int main(int argc, char **argv) { return db200h_cli(argc, argv); }
