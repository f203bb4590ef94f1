// Instantiation of the engine kernels for JW=1 (up to 32 jobs / transports per env).
#include "jsl_kernels.cuh"

namespace jsl {
JSL_DEFINE_JW(1)
}
