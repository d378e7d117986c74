// Host-side error plumbing shared by the C-ABI translation units.
#pragma once
#include <stdarg.h>
#include <stdio.h>

// Records a printf-style message for cn_last_error() (thread-local) and returns 1.
int cn_set_error(const char* fmt, ...);
