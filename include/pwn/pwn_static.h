// mirrors g2o_frontend/pwn_core/pwn_static.h -- the classes live in pwn/pwn.h
#pragma once
#include "pwn.h"
