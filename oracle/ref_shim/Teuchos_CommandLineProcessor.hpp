#pragma once
#include "Teuchos_shim.hpp"
