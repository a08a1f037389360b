"""TEST INFRASTRUCTURE ONLY -- empty pyplot stand-in."""
