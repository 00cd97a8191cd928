// TEST INFRASTRUCTURE: stand-in for <g2o/core/base_binary_edge.h>, see mini_g2o.h
#include "../mini_g2o.h"
