/* Test-infrastructure shim (NOT product code): stands in for gtest's
 * gtest_prod.h, which the reference's src/common.h:4 includes only for the
 * FRIEND_TEST macro (src/factor.h:89-91).  Lets the unmodified reference
 * sources compile without unpacking/building the vendored gtest zip. */
#ifndef DWG_ORACLE_SHIM_GTEST_PROD_H_
#define DWG_ORACLE_SHIM_GTEST_PROD_H_
#define FRIEND_TEST(test_case_name, test_name) \
  friend class test_case_name##_##test_name##_Test
#endif
