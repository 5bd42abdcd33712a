"""ORACLE support (test infrastructure): alias of the seeded synthetic state-dict generators.  They live in the product
package (nlc_b200.synthetic_weights) because bench.py's GPU arm builds its random-init networks from them and the product
arm must not import anything under oracle/; the oracle, the tests and the golden-vector generator keep this name."""
from nlc_b200.synthetic_weights import *  # noqa: F401,F403
from nlc_b200.synthetic_weights import ADM_CONFIGS, CONFIGS, EDM_CONFIGS  # noqa: F401
