#!/usr/bin/env python3
"""`python main.py` entry point, as in the reference's python-prototype/ (see upmix_b200/main.py)."""
from upmix_b200.main import main

if __name__ == "__main__":
    main()
