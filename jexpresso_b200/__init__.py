"""jexpresso_b200 -- B200-native explicit-RHS engine behind Jexpresso's rhs! surface."""
__version__ = "0.1.0"
