// TEST INFRASTRUCTURE: forwards to the Eigen stand-in (see ../../EigenShim.h).
#include "../../EigenShim.h"
