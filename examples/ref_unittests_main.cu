// Gate of the drop-in boundary: the reference's OWN unit tests -- every header its unittests/unittests.hxx includes
// (formats, graph, memory, framework/problem, operators for + advance merge_path, type limits, launch box, context,
// device properties, array container, .smtx loader, triangle counting) -- compiled UNCHANGED against this
// repository's include tree.  googletest is not in the image: tests/gtest_shim/gtest/gtest.h stands in for it
// (TEST / ASSERT_* / EXPECT_*, --gtest_filter, --gtest_list_tests).  Built by examples/build_reference_examples.sh
// with -I<reference>/unittests; the reference keeps its main() in gtest_main, this file is that main.
#include "unittests.hxx"

int main(int argc, char** argv) {
  ::testing::InitGoogleTest(&argc, argv);
  return RUN_ALL_TESTS();
}
