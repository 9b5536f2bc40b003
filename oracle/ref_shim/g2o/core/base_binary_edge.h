// TEST INFRASTRUCTURE ONLY: stand-in so that the reference header g2o_types/g2o_types.h compiles unmodified (see ../standin.h)
#pragma once
#include "../../standin.h"
