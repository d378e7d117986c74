from . import PATH
print(PATH)
