/* epoxy/glx.h -- TEST INFRASTRUCTURE (oracle).  Empty stand-in: horizonator-lib.c includes
 * it but uses nothing from it. */
#pragma once
