#pragma once
#include "Tpetra_Core.hpp"
