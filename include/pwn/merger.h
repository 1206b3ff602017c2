// mirrors g2o_frontend/pwn_core/merger.h -- the classes live in pwn/pwn.h
#pragma once
#include "pwn.h"
