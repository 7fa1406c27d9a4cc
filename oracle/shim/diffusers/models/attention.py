class BasicTransformerBlock:  # placeholder: never instantiated by the VC2 fixtures
    pass
