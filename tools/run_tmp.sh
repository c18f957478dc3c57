python tools/steplog.py 100000000 2>&1 | grep -E "^ +(15|16|17|18|19|20|3[6-9]|4[0-4]) "
