#!/bin/bash
( time python -m pytest tests -m gpu -x -q 2>&1 | tail -15 ) 2>&1
