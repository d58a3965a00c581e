// Source-compatibility shim: <ecsimd/utility.h> of aguinet/ecsimd, served by the B200 engine's mirror header.
#pragma once
#include "../../ecsimd.hpp"
