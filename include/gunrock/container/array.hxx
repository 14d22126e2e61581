/** @file array.hxx  Placeholder: the reference's std::array clone is off the hot path. */
#pragma once
#include <array>
