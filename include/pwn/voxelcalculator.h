// mirrors g2o_frontend/pwn_core/voxelcalculator.h -- the classes live in pwn/pwn.h
#pragma once
#include "pwn.h"
