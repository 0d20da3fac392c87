#pragma once
#define BASIX_VERSION 0.12.0.0
