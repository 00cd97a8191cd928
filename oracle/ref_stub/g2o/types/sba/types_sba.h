// TEST INFRASTRUCTURE: stand-in for <g2o/types/sba/types_sba.h>, see mini_g2o.h
#include "../../mini_g2o.h"
