"""Bench / smoke / test support (NOT part of the product package gill_b200): seeded synthetic stand-ins for the frozen
third-party models that cannot be downloaded offline, and the smoke check. Only this directory, tests/ and bench.py may
import oracle/."""
