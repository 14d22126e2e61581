/** @file cuda.hxx  Umbrella for the gcuda layer (include/gunrock/cuda/cuda.hxx). */
#pragma once
#include <gunrock/cuda/context.hxx>
#include <gunrock/util/math.hxx>
