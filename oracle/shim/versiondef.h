// Hand-written equivalent of the header CMake would configure from
// /root/reference/src/versiondef.h.in (project(lightmetrica VERSION 3.0.0)).
#pragma once
#define LM_VERSION_MAJOR 3
#define LM_VERSION_MINOR 0
#define LM_VERSION_PATCH 0
#define LM_VERSION_REVISION "70601db"
#define LM_BUILD_TIMESTAMP "oracle/_ref"
