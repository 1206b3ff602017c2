// mirrors g2o_frontend/pwn_core/multipointprojector.h -- the class lives in pwn/pwn.h
#pragma once
#include "pwn.h"
