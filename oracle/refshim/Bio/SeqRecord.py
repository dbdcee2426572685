from Bio.Seq import Seq


class SeqRecord:
    def __init__(self, seq, id="<unknown id>", name="<unknown name>",
                 description="<unknown description>"):
        self.seq = seq if isinstance(seq, Seq) else Seq(seq)
        self.id = id
        self.name = name
        self.description = description

    def __len__(self):
        return len(self.seq)

    def __iter__(self):
        return iter(self.seq)

    def __getitem__(self, index):
        if isinstance(index, slice):
            return SeqRecord(self.seq[index], id=self.id, name=self.name,
                             description=self.description)
        return self.seq[index]

    def format(self, fmt):
        if fmt != "fasta":
            raise ValueError(fmt)
        desc = self.description
        if desc and desc.split(None, 1)[:1] == [self.id]:
            title = desc
        elif desc and desc != "<unknown description>":
            title = f"{self.id} {desc}"
        else:
            title = self.id
        text = str(self.seq)
        body = "".join(text[i:i + 60] + "\n" for i in range(0, len(text), 60))
        return f">{title}\n{body}"
