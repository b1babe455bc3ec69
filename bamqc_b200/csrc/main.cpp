// bamqualcheck -- drop-in command for the reference CLI (src/bamqualcheck.cpp:239-457), statistics on the GPU.
#include "../../include/bamqc_b200.h"
int main(int argc, char** argv) { return bqc_main(argc, (const char* const*)argv); }
