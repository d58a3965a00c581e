#pragma once
#include "../shim.hpp"
