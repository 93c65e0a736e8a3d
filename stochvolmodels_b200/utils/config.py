"""Enumerations shared by the pricers; same names and values as the reference (utils/config.py:8-24)."""
from enum import Enum


class OptionType(str, Enum):
    CALL = "C"
    PUT = "P"
    INVERSE_CALL = "IC"
    INVERSE_PUT = "IP"


class VariableType(Enum):
    LOG_RETURN = 1   # transform variable PHI
    Q_VAR = 2        # transform variable PSI
    SIGMA = 3        # transform variable THETA
