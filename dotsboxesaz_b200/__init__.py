"""B200-native AlphaZero self-play engine for Dots & Boxes (drop-in for the Python surface of
damlobster/DotsBoxesAZ's self-play hot path).  See DESIGN.md and INTEGRATION.md."""
__version__ = "0.1.0"
