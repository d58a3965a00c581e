// Source-compatibility shim: <ecsimd/curve_group.h> of aguinet/ecsimd, served by the B200 engine's mirror header.
#pragma once
#include "../../ecsimd.hpp"
