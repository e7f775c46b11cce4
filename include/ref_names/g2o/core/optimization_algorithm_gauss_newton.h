// Forwarding header: the reference includes "g2o/core/optimization_algorithm_gauss_newton.h"; the B200 build resolves it to
// include/g2o_compat/g2o_compat.hpp. No declarations of its own.
#include "../../../g2o_compat/g2o_compat.hpp"
