// pointwise.h -- the parameter blocks live in the public header
#pragma once
#include "shiftgcn_b200.h"
