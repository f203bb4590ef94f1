// Instantiation of the engine kernels for JW=4 (up to 128 jobs / transports per env).
#include "jsl_kernels.cuh"

namespace jsl {
JSL_DEFINE_JW(4)
}
