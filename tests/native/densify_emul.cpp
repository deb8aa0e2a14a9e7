// TEST HARNESS (not product code): divshot_b200/csrc/densify.cu compiled for the host — the exact kernel bodies as
// serial loops and the exact host orchestration, CUDA runtime replaced by cuda_host_shim.h — so that the CPU suite
// (tests/test_densify_emul.py) can run the refinement step through the same dvs_densify_test_* hooks the staged GPU
// tests call in libgstrain.so.  The product never runs this.
#define DVS_DENSIFY_HOST_EMULATION 1
#include "../../divshot_b200/csrc/densify.cu"
