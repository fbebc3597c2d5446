rcParams = {}
