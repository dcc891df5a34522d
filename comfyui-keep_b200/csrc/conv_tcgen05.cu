// placeholder until the tcgen05 kernel lands
#include "ops.h"
using namespace keep;
int keepop_conv2d_tc(const ConvArgs& a, cudaStream_t s) { (void)a; (void)s; throw Error("tcgen05 conv not built yet"); }
