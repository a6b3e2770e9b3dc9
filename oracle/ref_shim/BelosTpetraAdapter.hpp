#pragma once
#include "Belos_shim.hpp"
