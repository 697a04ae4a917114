"""CPU oracle -- TEST INFRASTRUCTURE.  See oracle/jexref.c and oracle/ref.py headers."""
