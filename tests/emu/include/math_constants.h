// math_constants.h -- TEST INFRASTRUCTURE (see cuda_runtime.h in this directory)
#pragma once
#define CUDART_INF_F (__builtin_inff())
#define CUDART_NAN_F (__builtin_nanf(""))
#define CUDART_NAN (__builtin_nan(""))
