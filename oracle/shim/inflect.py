class engine:  # noqa: N801
    def plural(self, w, *a):
        return w + "s"

    def a(self, w):
        return "a " + w

    def number_to_words(self, n):
        return str(n)

    def singular_noun(self, w):
        return False
