/** @file operators.hxx  Umbrella (include/gunrock/framework/operators/operators.hxx:17-26). */
#pragma once
#include <gunrock/framework/operators/configs.hxx>
#include <gunrock/framework/operators/advance/advance.hxx>
#include <gunrock/framework/operators/filter/filter.hxx>
#include <gunrock/framework/operators/for/for.hxx>
#include <gunrock/framework/operators/uniquify/uniquify.hxx>
#include <gunrock/framework/operators/batch/batch.hxx>
#include <gunrock/framework/operators/neighborreduce/neighborreduce.hxx>
