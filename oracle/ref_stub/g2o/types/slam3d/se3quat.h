// TEST INFRASTRUCTURE: stand-in for <g2o/types/slam3d/se3quat.h>, see mini_g2o.h
#include "../../mini_g2o.h"
