/** @file framework.hxx  Umbrella (include/gunrock/framework/framework.hxx). */
#pragma once
#include <gunrock/framework/frontier/frontier.hxx>
#include <gunrock/framework/problem.hxx>
#include <gunrock/framework/enactor.hxx>
#include <gunrock/framework/benchmark.hxx>
#include <gunrock/framework/operators/operators.hxx>
