"""Reference-arm harness: stages and drives the UNMODIFIED reference (visionml/pytracking) -- never imported by pytracking_b200/."""
