// Source-compatibility shim: <ecsimd/jacobian_curve_point.h> of aguinet/ecsimd, served by the B200 engine's mirror header.
#pragma once
#include "../../ecsimd.hpp"
