/** @file cuda.hxx  Umbrella for the gcuda layer (include/gunrock/cuda/cuda.hxx). */
#pragma once
#include <gunrock/cuda/context.hxx>
#include <gunrock/cuda/launch_box.hxx>
#include <gunrock/util/math.hxx>
