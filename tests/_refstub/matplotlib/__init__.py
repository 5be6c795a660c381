"""Import stub: lets the reference's qpsim package import where matplotlib is absent (tests only)."""
