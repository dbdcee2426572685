def write(records, handle, fmt):
    count = 0
    for record in records:
        handle.write(record.format(fmt))
        count += 1
    return count
