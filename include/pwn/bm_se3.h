// mirrors g2o_frontend/pwn_core/bm_se3.h -- the classes live in pwn/pwn.h
#pragma once
#include "pwn.h"
