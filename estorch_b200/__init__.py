"""estorch_b200 -- Blackwell-native Evolution Strategies engine behind the estorch API."""
__version__ = "0.1.0"
