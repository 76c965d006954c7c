// Drop-in `count` program: the reference's count/count.cpp main(), served by the device path.
#include "../../../include/lesseq_b200.h"
int main(int argc, char **argv) { return lsqh_count_main(argc, argv); }
