"""Test-only binding of the CPU oracle (oracle/sk_oracle.c) through the same Python wrapper as the engine."""
import ctypes as C
import os
import subprocess

from skirt9_b200 import abi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_SO = os.path.join(ROOT, "oracle", "_build", "liboracle.so")


def build_oracle():
    src = os.path.join(ROOT, "oracle", "sk_oracle.c")
    if (not os.path.exists(ORACLE_SO)) or os.path.getmtime(ORACLE_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "oracle"])
    return ORACLE_SO


_lib = None


def oracle_library():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build_oracle())
    return _lib


def OracleEngine(config):
    return abi.Engine(config, lib=oracle_library(), prefix="sko_", misc_prefix="sko_")
