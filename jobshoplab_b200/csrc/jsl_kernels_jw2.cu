// Instantiation of the engine kernels for JW=2 (up to 64 jobs / transports per env).
#include "jsl_kernels.cuh"

namespace jsl {
JSL_DEFINE_JW(2)
}
