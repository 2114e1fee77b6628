/* raisr/RaisrVersion.h -- API level this library is a drop-in for (reference: Library/RaisrVersion.h:11-17). */
#ifndef RAISR_B200_VERSION_H
#define RAISR_B200_VERSION_H
#define RAISR_VERSION_MAJOR (23)
#define RAISR_VERSION_MINOR (11)
#define RAISR_CHECK_VERSION(major, minor) \
    (RAISR_VERSION_MAJOR > (major) || (RAISR_VERSION_MAJOR == (major) && RAISR_VERSION_MINOR >= (minor)))
#endif
