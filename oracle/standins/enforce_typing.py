"""Stand-in for `enforce-typing` (requirements.txt:4): identity decorator.
TEST INFRASTRUCTURE ONLY.  The real package only adds runtime type validation of
dataclass fields (config.py:75-270); it never changes a computed value."""


def enforce_types(wrapped):
    return wrapped
