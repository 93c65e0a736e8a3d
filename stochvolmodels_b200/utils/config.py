"""Enumerations shared by the pricers: the names and values callers of the reference already use (utils/config.py:8-24)."""
from enum import Enum

# payoff codes as they appear in ``optiontypes_ttms``; a str-mixin Enum, so ``OptionType.CALL == "C"`` holds
OptionType = Enum("OptionType", {"CALL": "C", "PUT": "P", "INVERSE_CALL": "IC", "INVERSE_PUT": "IP"}, type=str, module=__name__)

# which state variable a transform / payoff refers to: log-return (grid PHI), quadratic variance (grid PSI), volatility (grid THETA)
VariableType = Enum("VariableType", {"LOG_RETURN": 1, "Q_VAR": 2, "SIGMA": 3}, module=__name__)
