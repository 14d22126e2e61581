#pragma once
#include <gunrock/graph/graph.hxx>
