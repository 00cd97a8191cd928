// TEST INFRASTRUCTURE: stand-in for <g2o/core/eigen_types.h>, see mini_g2o.h
#include "../mini_g2o.h"
