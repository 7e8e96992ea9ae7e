/* util.h -- diagnostic convention shared with the reference (reference util.h:4):
 * one line on stderr, "file(line) at func(): message". */
#pragma once
#include <stdio.h>
#define MSG(fmt, ...) \
    fprintf(stderr, "%s(%d) at %s(): " fmt "\n", __FILE__, __LINE__, __func__, ##__VA_ARGS__)
