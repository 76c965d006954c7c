// Drop-in `solve` program: the reference's solve/solve.cpp main(), served by the device path.
#include "../../../include/lesseq_b200.h"
int main(int argc, char **argv) { return lsqh_solve_main(argc, argv); }
