// mirrors g2o_frontend/pwn_core/linearizer.h -- the classes live in pwn/pwn.h
#pragma once
#include "pwn.h"
