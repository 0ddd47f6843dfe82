// NTL/ZZ.h stand-in — see RR.h in this directory (TEST INFRASTRUCTURE ONLY).
#pragma once
#include "RR.h"
