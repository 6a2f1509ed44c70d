#pragma once
#include "defines.hpp"
#include "lbm.hpp"
#include "shapes.hpp"
void main_setup(); // the scene; one definition is active in setup.cpp
