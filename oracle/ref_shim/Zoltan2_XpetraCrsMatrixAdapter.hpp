#pragma once
#include "Zoltan2_Adapter.hpp"
