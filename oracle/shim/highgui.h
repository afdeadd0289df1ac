// oracle/shim/highgui.h -- TEST INFRASTRUCTURE. See cv.h in this directory.
#include "cv.h"
