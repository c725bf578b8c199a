// cuda_pt — drop-in command line of the reference (src/main.cpp:9-25); all logic
// lives in libb200pt.so (pt_cli_main).
#include "../../include/b200pt.h"
int main(int argc, char** argv) { return pt_cli_main(argc, argv); }
