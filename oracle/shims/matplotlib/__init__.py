"""ORACLE / TEST INFRASTRUCTURE: empty stand-in so `import matplotlib.pyplot` succeeds."""
