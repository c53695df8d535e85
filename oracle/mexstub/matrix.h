// TEST INFRASTRUCTURE ONLY: see mex.h in this directory.
#include "mex.h"
