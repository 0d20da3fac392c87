"""``basix.sobolev_spaces`` (cpp/basix/sobolev-spaces.h:12-22)."""

from enum import IntEnum


class SobolevSpace(IntEnum):
    L2 = 0
    H1 = 1
    H2 = 2
    H3 = 3
    HInf = 8
    HDiv = 10
    HCurl = 11
    HEin = 12
    HDivDiv = 13
