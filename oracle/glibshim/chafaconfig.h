/* chafaconfig.h -- stand-in for the file the reference generates from
 * chafa/chafaconfig.h.in (version from configure.ac:7-11). TEST INFRASTRUCTURE ONLY. */
#ifndef __CHAFACONFIG_H__
#define __CHAFACONFIG_H__
#define CHAFA_MAJOR_VERSION 1
#define CHAFA_MINOR_VERSION 19
#define CHAFA_MICRO_VERSION 0
#endif
