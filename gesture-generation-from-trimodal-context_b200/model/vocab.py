"""Word / speaker index dictionary - the boundary type PoseGenerator inspects (reference: scripts/model/vocab.py:8-37;
`isinstance(z_obj, vocab.Vocab)` at multimodal_context_net.py:87, `z_obj.n_words` at :89).  Only the indexing surface
is provided; fastText loading (vocab.py:70-130) is data tooling outside the hot path."""


class Vocab:
    PAD_token, SOS_token, EOS_token, UNK_token = 0, 1, 2, 3

    def __init__(self, name, insert_default_tokens=True):
        self.name = name
        self.trimmed = False
        self.word_embedding_weights = None
        self.reset_dictionary(insert_default_tokens)

    def reset_dictionary(self, insert_default_tokens=True):
        self.word2index, self.word2count = {}, {}
        if insert_default_tokens:
            self.index2word = {self.PAD_token: '<PAD>', self.SOS_token: '<SOS>', self.EOS_token: '<EOS>', self.UNK_token: '<UNK>'}
        else:
            self.index2word = {self.UNK_token: '<UNK>'}
        self.n_words = len(self.index2word)

    def index_word(self, word):
        if word in self.word2index:
            self.word2count[word] += 1
            return
        self.word2index[word] = self.n_words
        self.word2count[word] = 1
        self.index2word[self.n_words] = word
        self.n_words += 1

    def add_vocab(self, other_vocab):
        for word in other_vocab.word2count:
            self.index_word(word)

    def get_word_index(self, word):
        return self.word2index.get(word, self.UNK_token)
