#pragma once
#include "../string.hpp"
